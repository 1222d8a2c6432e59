"""Mirror of utils_criterion.py (reference :6-90) on the GPU.

`compute_errors(gt, pred)` keeps the reference's signature and return order
(abs_rel, rmse, a1, a2, a3, log_10, mae) for one depth map; `batch_errors(gt, pred, cfg)` is the per-sample loop of
the validation / test code around it (train.py:795-838, test.py:231-285: x max_depth when depth_norm, clip the
prediction to [eps, max_depth], gt >= 0) for a whole batch in one library call, returning a [B,7] float64 device
tensor -- no per-sample device -> host copies; take `.mean(0)` for the epoch numbers the reference prints.
"""
import torch

from . import _lib

METRIC_NAMES = ("abs_rel", "rmse", "delta1", "delta2", "delta3", "log10", "mae")


def _run(gt, pred, batch, scale, prepare, lo, hi):
    _lib.require_cuda(pred, "pred", torch.float32)
    _lib.require_cuda(gt, "gt", torch.float32)
    if gt.numel() != pred.numel():
        raise ValueError("gt and pred differ in size: %s vs %s" % (tuple(gt.shape), tuple(pred.shape)))
    pred, gt = pred.contiguous(), gt.contiguous()
    lib = _lib.load()
    ws = torch.empty(lib.adp_depth_metrics_workspace_bytes(batch), dtype=torch.uint8, device=pred.device)
    out = torch.empty((batch, 7), dtype=torch.float64, device=pred.device)
    with _lib.on_device(pred):
        _lib.check(lib.adp_depth_metrics(pred.data_ptr(), gt.data_ptr(), batch, pred.numel() // batch, float(scale),
                                         int(prepare), float(lo), float(hi), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _lib.stream_ptr()))
    return out


def batch_errors(gt, pred, cfg=None, depth_norm=None, max_depth=None, protocol="train"):
    """gt, pred: [B,1,H,W] CUDA fp32 (network output and loader ground truth, un-denormalised).  -> [B,7] float64.
    protocol 'train': the validation pass of train.py:807-825 (pred clipped to [eps, max_depth]);
    protocol 'test' : test.py:262-276 (pred and gt only clipped at 0)."""
    if cfg is not None:
        depth_norm, max_depth = bool(cfg.dataset.depth_norm), float(cfg.dataset.max_depth)
    scale = max_depth if depth_norm else 1.0
    if protocol == "test":
        return _run(gt, pred, pred.shape[0], scale, 1, 0.0, float("inf"))
    if protocol != "train":
        raise ValueError("protocol must be 'train' or 'test', got %r" % (protocol,))
    eps = 1e-3 if depth_norm else 1e-6                       # train.py:823
    return _run(gt, pred, pred.shape[0], scale, 1, eps, max_depth)


def compute_errors(gt, pred, min_depth_threshold=0.0):
    """One depth map (any shape), already in metres: the reference function itself.  Returns 7 Python floats
    (this synchronises; use batch_errors inside loops).  `min_depth_threshold` is accepted and ignored exactly as
    in the reference (:22-24)."""
    if not torch.is_tensor(gt):
        gt = torch.as_tensor(gt, dtype=torch.float32, device="cuda")
    if not torch.is_tensor(pred):
        pred = torch.as_tensor(pred, dtype=torch.float32, device=gt.device)
    return tuple(float(v) for v in _run(gt.float(), pred.float(), 1, 1.0, 0, 0.0, 0.0)[0].tolist())
