"""B200 mirror of utils_loss.py (reference :9-49) and of the criterion wiring of the training loop
(train.py:420-467, :646-669).

`SIlogLoss(lambda_scale, epsilon)(pred, target)` keeps the reference signature (already-masked
tensors of any shape -> 0-dim tensor).  `DepthCriterion` is the whole of train.py:646-669 -- mask
`gt != 0`, optional x max_depth, L1 / SIlog / Combined -- as one fused statistics kernel and one
fused gradient kernel (adp_depth_loss_sums / _value / _backward): no boolean-index gather, no
`.item()` synchronisation.  `reduce_fn` lets a data-parallel trainer all-reduce the four sufficient
statistics so that the loss is the global-batch loss the reference computes after DataParallel
gathers the predictions on GPU 0.
"""
import torch
import torch.nn as nn

from . import _lib


class _DepthLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, scale, eps, use_mask, l1_w, silog_w, lam, reduce_fn):
        lib = _lib.load()
        _lib.require_cuda(pred, "pred", torch.float32)
        _lib.require_cuda(gt, "target", torch.float32)
        if pred.shape != gt.shape:
            raise ValueError("pred and target shapes differ: %s vs %s" % (tuple(pred.shape), tuple(gt.shape)))
        pred, gt = pred.contiguous(), gt.contiguous()
        sums = torch.zeros(4, device=pred.device, dtype=torch.float64)
        out = torch.empty(3, device=pred.device, dtype=torch.float32)
        with torch.cuda.device(pred.device):
            s = _lib.stream_ptr()
            _lib.check(lib.adp_depth_loss_sums(pred.data_ptr(), gt.data_ptr(), pred.numel(), scale, eps, use_mask,
                                               sums.data_ptr(), s))
            if reduce_fn is not None:
                reduce_fn(sums)
            _lib.check(lib.adp_depth_loss_value(sums.data_ptr(), l1_w, silog_w, lam, out.data_ptr(), s))
        ctx.save_for_backward(pred, gt, sums)
        ctx.cfg = (scale, eps, use_mask, l1_w, silog_w, lam)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, grad_loss, _grad_parts):
        pred, gt, sums = ctx.saved_tensors
        scale, eps, use_mask, l1_w, silog_w, lam = ctx.cfg
        lib = _lib.load()
        dpred = torch.empty_like(pred)
        gscale = grad_loss.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(pred.device):
            _lib.check(lib.adp_depth_loss_backward(pred.data_ptr(), gt.data_ptr(), pred.numel(), scale, eps,
                                                   use_mask, sums.data_ptr(), l1_w, silog_w, lam,
                                                   gscale.data_ptr(), dpred.data_ptr(), _lib.stream_ptr()))
        return dpred, None, None, None, None, None, None, None, None


class SIlogLoss(nn.Module):
    """Scale-invariant logarithmic loss, reference utils_loss.py:9-49:
    sqrt(clamp(mean(d^2) - lambda*mean(d)^2, 0)),  d = log(clamp(pred,eps)) - log(clamp(target,eps))."""

    def __init__(self, lambda_scale=0.5, epsilon=1e-6):
        super().__init__()
        self.lambda_scale = lambda_scale
        self.epsilon = epsilon

    def forward(self, pred, target):
        loss, _ = _DepthLossFunction.apply(pred, target, 1.0, float(self.epsilon), 0, 0.0, 1.0,
                                           float(self.lambda_scale), None)
        return loss


class DepthCriterion(nn.Module):
    """train.py:420-467 (criterion selection) + :646-669 (mask, denormalisation, weighting).

    criterion in {'L1', 'SIlog', 'Combined'}.  forward(depth_pred, gtdepth) takes the UNMASKED
    [B,1,S,S] tensors and returns the scalar loss; `.last_parts` holds (loss, l1, silog) on device.
    """

    def __init__(self, criterion="Combined", l1_weight=0.237, silog_weight=0.637, silog_lambda=0.869,
                 depth_norm=False, max_depth=30.0, epsilon=1e-6, reduce_fn=None):
        super().__init__()
        if criterion == "L1":
            self.l1_w, self.silog_w = 1.0, 0.0
        elif criterion == "SIlog":
            self.l1_w, self.silog_w = 0.0, 1.0
        elif criterion == "Combined":
            self.l1_w, self.silog_w = float(l1_weight), float(silog_weight)
            if self.silog_w == 0.0:            # train.py:446-459: SIlog disabled -> "L1 only" with weight 1.0
                self.l1_w = 1.0
        else:
            raise ValueError("Unknown criterion: %s" % criterion)
        self.criterion = criterion
        self.lam = float(silog_lambda)
        self.scale = float(max_depth) if depth_norm else 1.0
        self.eps = float(epsilon)
        self.reduce_fn = reduce_fn
        self.last_parts = None

    @classmethod
    def from_cfg(cls, cfg, reduce_fn=None):
        m = cfg.mode
        return cls(getattr(m, "criterion", "L1"), getattr(m, "l1_weight", 0.5), getattr(m, "silog_weight", 0.5),
                   getattr(m, "silog_lambda", 0.5), bool(cfg.dataset.depth_norm), cfg.dataset.max_depth,
                   reduce_fn=reduce_fn)

    def forward(self, depth_pred, gtdepth):
        loss, parts = _DepthLossFunction.apply(depth_pred, gtdepth, self.scale, self.eps, 1, self.l1_w,
                                               self.silog_w, self.lam, self.reduce_fn)
        self.last_parts = parts
        return loss
