"""numpy (float64) restatement of the masked depth loss and its gradient.

Test infrastructure only (see oracle/__init__.py).  Parity: pinned against
utils_loss.SIlogLoss + nn.L1Loss + autograd of the unmodified reference by
tests/test_oracle_golden.py.

Follows train.py:646-669 (mask gt != 0, optional x max_depth, Combined =
l1_w*L1 + silog_w*SIlog over the valid pixels of the whole batch) and
utils_loss.py:29-49 (clamp(min=eps), d = log p - log g,
sqrt(clamp(mean(d^2) - lambda*mean(d)^2, 0))).
"""
import numpy as np


def loss_sums(pred, gt, scale=1.0, eps=1e-6):
    """The four sufficient statistics (N_valid, sum|p-g|, sum d, sum d^2)."""
    p = pred.astype(np.float64).ravel() * scale
    g = gt.astype(np.float64).ravel() * scale
    m = gt.ravel() != 0.0
    p, g = p[m], g[m]
    d = np.log(np.maximum(p, eps)) - np.log(np.maximum(g, eps))
    return float(m.sum()), float(np.abs(p - g).sum()), float(d.sum()), float((d * d).sum())


def loss_from_sums(n, s_abs, s_d, s_d2, l1_w, silog_w, lam):
    l1 = s_abs / n
    m1, m2 = s_d / n, s_d2 / n
    v = m2 - lam * m1 * m1
    si = np.sqrt(max(v, 0.0))
    return l1_w * l1 + silog_w * si, l1, si


def depth_loss_and_grad(pred, gt, l1_w=0.237, silog_w=0.637, lam=0.869, scale=1.0, eps=1e-6):
    """Returns (loss, l1, silog, dloss/dpred) -- SURVEY.md App. B.6 closed form."""
    n, s_abs, s_d, s_d2 = loss_sums(pred, gt, scale, eps)
    loss, l1, si = loss_from_sums(n, s_abs, s_d, s_d2, l1_w, silog_w, lam)
    p = pred.astype(np.float64) * scale
    g = gt.astype(np.float64) * scale
    m = gt != 0.0
    grad = np.zeros_like(p)
    grad += l1_w * scale * np.sign(p - g) / n
    if silog_w != 0.0:
        pc = np.maximum(p, eps)
        d = np.log(pc) - np.log(np.maximum(g, eps))
        m1 = s_d / n
        with np.errstate(divide="ignore", invalid="ignore"):
            gs = scale * (p >= eps) * (d - lam * m1) / (n * si * pc)
        grad += silog_w * gs
    grad = np.where(m, grad, 0.0)
    return loss, l1, si, grad
