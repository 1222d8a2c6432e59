"""numpy restatement of the reference's evaluation metrics.

Test infrastructure only (see oracle/__init__.py).  Parity: pinned against utils_criterion.compute_errors of the
unmodified reference by tests/test_oracle_golden.py (tests/golden/metrics.npz).

Follows utils_criterion.py:6-90 (mask gt != 0; adaptive epsilon 1e-3 if max(gt) > 1 else 1e-6; valid = pred > eps &
gt > eps with the fall-backs of :38-54; thresholds 1.25^k; rmse, abs_rel, log10, mae; NaN/inf -> 0) and the
per-sample preparation of the validation loop, train.py:807-825 (x max_depth when depth_norm, clip pred to
[eps, max_depth], gt >= 0).
"""
import numpy as np


def prepare(gt, pred, depth_norm, max_depth, protocol="train"):
    """train.py:807-825 (validation inside the training script: pred clipped to [eps, max_depth]) or
    test.py:262-276 (protocol='test': both maps only clipped at 0)."""
    gt = gt.astype(np.float32)
    pred = pred.astype(np.float32)
    if depth_norm:
        gt = gt * np.float32(max_depth)
        pred = pred * np.float32(max_depth)
    if protocol == "test":
        return np.maximum(gt, 0.0), np.maximum(pred, 0.0)
    eps = np.float32(1e-3 if depth_norm else 1e-6)
    return np.maximum(gt, 0.0), np.clip(pred, eps, np.float32(max_depth))


def compute_errors(gt, pred):
    """(abs_rel, rmse, a1, a2, a3, log_10, mae) of one depth map -- utils_criterion.py:6-90.
    Element-wise arithmetic and comparisons in float32 with float32 epsilons, as numpy does for the float32 maps
    the reference passes (a prediction clipped to float32(1e-3) is NOT > epsilon); means accumulate in float64."""
    f32 = np.float32
    gt = np.asarray(gt, dtype=f32).ravel()
    pred = np.asarray(pred, dtype=f32).ravel()
    mask = gt != 0.0
    if mask.sum() == 0:
        return (0.0,) * 7
    pred, gt = pred[mask], gt[mask]
    eps = f32(1e-3) if gt.max() > 1.0 else f32(1e-6)
    valid = (pred > eps) & (gt > eps)
    if valid.sum() == 0:
        valid = gt > eps
        if valid.sum() == 0:
            return (0.0,) * 7
        valid = valid & (pred > 0)
        if valid.sum() == 0:
            return 1.0, float(gt.max()), 0.0, 0.0, 0.0, 1.0, float(gt.max())
    pred, gt = pred[valid], gt[valid]
    eps = f32(1e-3) if gt.max() > 1.0 else f32(1e-6)
    pc = np.maximum(pred, eps)
    thresh = np.maximum(gt / pc, pc / gt)
    mean = lambda a: float(np.mean(a, dtype=np.float64))
    vals = [mean(np.abs(gt - pred) / gt), np.sqrt(mean((gt - pred) ** 2)), mean(thresh < f32(1.25)),
            mean(thresh < f32(1.25 ** 2)), mean(thresh < f32(1.25 ** 3)),
            mean(np.abs(np.log10(np.maximum(gt, eps)) - np.log10(pc))), mean(np.abs(gt - pred))]
    return tuple(0.0 if (v != v or v == np.inf) else float(v) for v in vals)


def batch_errors(gt, pred, depth_norm, max_depth, protocol="train"):
    """[B,1,H,W] -> [B,7]: the per-sample loop of train.py:795-838 / test.py:243-276."""
    out = []
    for i in range(gt.shape[0]):
        g, p = prepare(gt[i], pred[i], depth_norm, max_depth, protocol)
        out.append(compute_errors(g, p))
    return np.array(out, dtype=np.float64)
