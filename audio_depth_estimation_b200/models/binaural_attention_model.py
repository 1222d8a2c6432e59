"""Mirror of models/binaural_attention_model.py (reference :22-340) -- BASELINE config 4 -- on libadp_b200.

Same classes, constructor arguments and `state_dict` (keys and shapes) as the reference, so its checkpoints load:
`DoubleConv`, `Down`, `Up`, `BinauralCrossAttention`, `BinauralEncoder`, `BinauralAttentionDepthNet`.  The nn.Module
tree only HOLDS the parameters; the arithmetic runs in `BinauralAttentionDepthNet.forward` on bf16 NHWC activations
through the C ABI:

  3x3 convolutions (:29,:32)        adp_conv2d_k3s1_{fprop,dgrad,wgrad} (tcgen05 implicit GEMM, the U-Net's kernel),
                                    adp_conv2d_k3s1_c1_* for the one-channel stem
  BatchNorm + ReLU (:30-34)         adp_bn_act_{forward,backward}
  MaxPool2d / Upsample (:48,:62)    adp_maxpool2_*, adp_upsample2x_*; torch.cat (:75,:306) is never materialised
  1x1 convolutions (:97-103,:241)   adp_gemm_rows_bf16 / adp_gemm_tn_bf16 (+ adp_rows_op / adp_rows_reduce for biases)
  cross attention (:106-153)        adp_gemm_rows_softmax (score GEMM with the softmax / softmax-backward in its epilogue)
                                    + adp_gemm_rows_bf16: the [HW x HW] matrices exist only in bf16, one sample and
                                    direction at a time; the backward pass recomputes them in both orientations
                                    instead of transposing
  head (:262-265, :318-332)         adp_depth_head_{forward,backward}

  Up, bilinear=False (:65-66)       adp_gemm_rows_bf16 (4N columns per input pixel) + adp_pixel_shuffle2
  output resize (:322-328)          adp_bilinear_resize_{forward,backward} (F.interpolate, align_corners=False)

Restrictions: bf16 only, H = W = a power of two with H/16 >= 8 (every attention level needs a multiple of 64 tokens; the
F.pad of :72-73 is then a no-op), base_channels a multiple of 64.
"""
import math

import torch
import torch.nn as nn
from torch.nn import init

from .. import _lib

_BF16 = torch.bfloat16


def _sp():
    return _lib.stream_ptr()


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _scratch_for(rows, n, device):
    """fp32 split-K scratch for small grids (the kernel only splits when fewer tiles than SMs exist)."""
    if rows * n * 4 <= (32 << 20) and rows < 148 * 128:
        return torch.empty(rows * n, device=device, dtype=torch.float32)
    return None


def _w16(weight, pad_rows=0):
    """bf16 copy of a conv weight in its channels_last memory order [Cout][kh][kw][Cin] (optionally zero-padded to
    `pad_rows` output channels)."""
    cout = weight.shape[0]
    n = weight.numel()
    per = n // cout
    out = torch.zeros(max(cout, pad_rows) * per, device=weight.device, dtype=_BF16) if pad_rows > cout else \
        torch.empty(n, device=weight.device, dtype=_BF16)
    _lib.check(_lib.load().adp_cast_bf16(weight.data_ptr(), out.data_ptr(), n, _sp()))
    return out


def _channels_last_(conv):
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)


# ============================================================================================ autograd functions
class _Conv3x3(torch.autograd.Function):
    """x0 (| x1): bf16 [B,H,W,C] -> bf16 [B,H,W,Cout]"""

    @staticmethod
    def forward(ctx, x0, x1, weight):
        B, H, W, C0 = x0.shape
        C1 = x1.shape[-1] if x1 is not None else 0
        cout = weight.shape[0]
        w16 = _w16(weight)
        y = torch.empty((B, H, W, cout), device=x0.device, dtype=_BF16)
        sc = _scratch_for(B * H * W, cout, x0.device)
        _lib.check(_lib.load().adp_conv2d_k3s1_fprop(x0.data_ptr(), C0, _ptr(x1), C1, w16.data_ptr(), y.data_ptr(), B, H, W, cout,
                                                     _ptr(sc), sc.numel() * 4 if sc is not None else 0, _sp()))
        ctx.save_for_backward(x0, x1, weight, w16)
        return y

    @staticmethod
    def backward(ctx, dy):
        x0, x1, weight, w16 = ctx.saved_tensors
        B, H, W, C0 = x0.shape
        C1 = x1.shape[-1] if x1 is not None else 0
        cout = weight.shape[0]
        dy = dy.contiguous()
        lib = _lib.load()
        dx0 = torch.empty_like(x0)
        dx1 = torch.empty_like(x1) if x1 is not None else None
        sc = _scratch_for(B * H * W, C0 + C1, x0.device)
        _lib.check(lib.adp_conv2d_k3s1_dgrad(dy.data_ptr(), cout, w16.data_ptr(), dx0.data_ptr(), C0, _ptr(dx1), C1, B, H, W,
                                             _ptr(sc), sc.numel() * 4 if sc is not None else 0, _sp()))
        dw = torch.empty_like(weight)            # channels_last strides: memory [Cout][3][3][Cin]
        _lib.check(lib.adp_conv2d_k3s1_wgrad(dy.data_ptr(), cout, x0.data_ptr(), C0, _ptr(x1), C1, dw.data_ptr(), B, H, W, _sp()))
        return dx0, dx1, dw


class _Conv3x3Stem(torch.autograd.Function):
    """one plane of the fp32 [B,2,H,W] network input -> bf16 [B,H,W,Cout]"""

    @staticmethod
    def forward(ctx, x, channel, weight):
        B, _, H, W = x.shape
        cout = weight.shape[0]
        y = torch.empty((B, H, W, cout), device=x.device, dtype=_BF16)
        plane = x.data_ptr() + 4 * channel * H * W
        _lib.check(_lib.load().adp_conv2d_k3s1_c1_fprop(plane, x.stride(0), weight.data_ptr(), y.data_ptr(), B, H, W, cout, _sp()))
        ctx.save_for_backward(x, weight)
        ctx.channel = channel
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, _, H, W = x.shape
        dy = dy.contiguous()
        dw = torch.empty_like(weight)
        plane = x.data_ptr() + 4 * ctx.channel * H * W
        _lib.check(_lib.load().adp_conv2d_k3s1_c1_wgrad(dy.data_ptr(), plane, x.stride(0), dw.data_ptr(), B, H, W, weight.shape[0], _sp()))
        return None, None, dw


class _BnRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, conv_bias, bn, training):
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        saved = torch.empty(4 * C, device=x.device, dtype=torch.float32)
        ws = torch.empty(2 * C, device=x.device, dtype=torch.float64)
        mom = bn.momentum if bn.momentum is not None else 0.1
        _lib.check(_lib.load().adp_bn_act_forward(x.data_ptr(), rows, C, gamma.data_ptr(), beta.data_ptr(), _ptr(conv_bias),
                                                  bn.running_mean.data_ptr(), bn.running_var.data_ptr(), int(training), bn.eps, mom,
                                                  0.0, y.data_ptr(), saved.data_ptr(), ws.data_ptr(), _sp()))
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(x, saved)
        ctx.training = bool(training)
        ctx.has_bias = conv_bias is not None
        ctx.bias_like = conv_bias
        return y

    @staticmethod
    def backward(ctx, dy):
        x, saved = ctx.saved_tensors
        C = x.shape[-1]
        rows = x.numel() // C
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg = torch.empty(C, device=x.device, dtype=torch.float32)
        db = torch.empty(C, device=x.device, dtype=torch.float32)
        ws = torch.empty(2 * C, device=x.device, dtype=torch.float64)
        _lib.check(_lib.load().adp_bn_act_backward(x.data_ptr(), rows, C, saved.data_ptr(), dy.data_ptr(), 0.0, int(ctx.training),
                                                   dx.data_ptr(), dg.data_ptr(), db.data_ptr(), ws.data_ptr(), _sp()))
        # a bias in front of a batch-statistics BatchNorm has zero gradient (the normalisation removes it)
        dbias = None
        if ctx.has_bias:
            if ctx.training:
                dbias = torch.zeros_like(ctx.bias_like)
            else:
                dbias = torch.empty_like(ctx.bias_like)
                _lib.check(_lib.load().adp_rows_reduce(0, dx.data_ptr(), None, rows, C, dbias.data_ptr(), ws.data_ptr(), _sp()))
        return dx, dg, db, dbias, None, None


class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, H, W, C = x.shape
        y = torch.empty((B, H // 2, W // 2, C), device=x.device, dtype=_BF16)
        _lib.check(_lib.load().adp_maxpool2_forward(x.data_ptr(), y.data_ptr(), B, H // 2, W // 2, C, _sp()))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, H, W, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        _lib.check(_lib.load().adp_maxpool2_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, H // 2, W // 2, C, _sp()))
        return dx


class _Upsample2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, H, W, C = x.shape
        y = torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=_BF16)
        _lib.check(_lib.load().adp_upsample2x_forward(x.data_ptr(), y.data_ptr(), B, H, W, C, _sp()))
        ctx.shape = (B, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty((B, H, W, C), device=dy.device, dtype=_BF16)
        _lib.check(_lib.load().adp_upsample2x_backward(dy.data_ptr(), dx.data_ptr(), B, H, W, C, _sp()))
        return dx


class _Conv1x1(torch.autograd.Function):
    """(x0 | x1) [.., K] @ weight[N][K]^T (+ bias) -> [.., Np]; Np = N rounded up to 64 (extra columns are zero)."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias):
        K0 = x0.shape[-1]
        K1 = x1.shape[-1] if x1 is not None else 0
        N = weight.shape[0]
        Np = (N + 63) // 64 * 64
        rows = x0.numel() // K0
        w16 = _w16(weight, pad_rows=Np)
        y = torch.empty(x0.shape[:-1] + (Np,), device=x0.device, dtype=_BF16)
        lib = _lib.load()
        _lib.check(lib.adp_gemm_rows_bf16(x0.data_ptr(), K0, _ptr(x1), K1, w16.data_ptr(), 0, y.data_ptr(), Np, None, 0, None, rows, _sp()))
        if bias is not None:
            b = bias if Np == N else torch.cat([bias, bias.new_zeros(Np - N)])
            _lib.check(lib.adp_rows_op(0, y.data_ptr(), None, b.data_ptr(), y.data_ptr(), rows, Np, _sp()))
        ctx.save_for_backward(x0, x1, weight, w16)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x0, x1, weight, w16 = ctx.saved_tensors
        K0 = x0.shape[-1]
        K1 = x1.shape[-1] if x1 is not None else 0
        N = weight.shape[0]
        Np = dy.shape[-1]
        rows = x0.numel() // K0
        dy = dy.contiguous()
        lib = _lib.load()
        dx0 = torch.empty_like(x0)
        dx1 = torch.empty_like(x1) if x1 is not None else None
        # dx = dy [rows][Np] @ w16 [Np][K]   ("NN": the weight is the [K' = Np][N' = K] operand)
        _lib.check(lib.adp_gemm_rows_bf16(dy.data_ptr(), Np, None, 0, w16.data_ptr(), 1, dx0.data_ptr(), K0, _ptr(dx1), K1, None, rows, _sp()))
        dwp = torch.zeros((Np, K0 + K1), device=dy.device, dtype=torch.float32)
        _lib.check(lib.adp_gemm_tn_bf16(dy.data_ptr(), Np, x0.data_ptr(), K0, dwp.data_ptr(), K0 + K1, rows, _sp()))
        if x1 is not None:
            _lib.check(lib.adp_gemm_tn_bf16(dy.data_ptr(), Np, x1.data_ptr(), K1, dwp.data_ptr() + 4 * K0, K0 + K1, rows, _sp()))
        dw = dwp[:N].reshape(weight.shape)
        db = None
        if ctx.has_bias:
            dbp = torch.empty(Np, device=dy.device, dtype=torch.float32)
            ws = torch.empty(2 * Np, device=dy.device, dtype=torch.float64)
            _lib.check(lib.adp_rows_reduce(0, dy.data_ptr(), None, rows, Np, dbp.data_ptr(), ws.data_ptr(), _sp()))
            db = dbp[:N]
        return dx0, dx1, dw, db


class _Attend(torch.autograd.Function):
    """softmax(Q K^T / sqrt(C)) V per sample (:118-127).  q, k: [B,T,64p] (zero-padded projections), v: [B,T,C].

    The [T x T] matrices exist only in bf16: the score GEMM is run three times in the forward pass (row max, row sum,
    probabilities -- its K dimension is 64, so it is cheap next to P V) with the softmax in its epilogue; the backward
    pass recomputes P the same way, folds dS = scale * P * (dP - delta) into the epilogue of the dP GEMM, and forms the
    transposed products dV = P^T dO, dK = dS^T Q with the MN-major x MN-major GEMM of the weight gradients."""

    @staticmethod
    def forward(ctx, q, k, v, scale):
        B, T, Dq = q.shape
        C = v.shape[-1]
        lib = _lib.load()
        o = torch.empty_like(v)
        m = torch.empty((B, T), device=q.device, dtype=torch.int32)       # order-preserving int image of the row max
        l = torch.empty((B, T), device=q.device, dtype=torch.float32)
        P = torch.empty((T, T), device=q.device, dtype=_BF16)
        _lib.check(lib.adp_softmax_stats_init(m.data_ptr(), l.data_ptr(), B * T, _sp()))
        for b in range(B):
            for mode in (1, 2, 3):
                _lib.check(lib.adp_gemm_rows_softmax(q[b].data_ptr(), Dq, k[b].data_ptr(), T, T, mode, 0, scale, m[b].data_ptr(),
                                                     l[b].data_ptr(), None, None, P.data_ptr() if mode == 3 else None, _sp()))
            _lib.check(lib.adp_gemm_rows_bf16(P.data_ptr(), T, None, 0, v[b].data_ptr(), 1, o[b].data_ptr(), C, None, 0, None, T, _sp()))
        ctx.save_for_backward(q, k, v, o, m, l)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, m, l = ctx.saved_tensors
        B, T, Dq = q.shape
        C = v.shape[-1]
        scale = ctx.scale
        do = do.contiguous()
        lib = _lib.load()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dev = q.device
        P = torch.empty((T, T), device=dev, dtype=_BF16)
        D = torch.empty((T, T), device=dev, dtype=_BF16)
        delta = torch.empty(T, device=dev, dtype=torch.float32)

        dvf = torch.empty((T, C), device=dev, dtype=torch.float32)
        dkf = torch.empty((T, Dq), device=dev, dtype=torch.float32)

        def tn(a, bmat, n, acc, out):       # out[j][n] = sum_i a[i][j] * bmat[i][n]: both operands row-major, no transpose
            acc.zero_()
            _lib.check(lib.adp_gemm_tn_bf16(a.data_ptr(), T, bmat.data_ptr(), n, acc.data_ptr(), n, T, _sp()))
            _lib.check(lib.adp_cast_bf16(acc.data_ptr(), out.data_ptr(), T * n, _sp()))

        for b in range(B):
            _lib.check(lib.adp_rows_reduce(1, do[b].data_ptr(), o[b].data_ptr(), T, C, delta.data_ptr(), None, _sp()))
            # P [i][j] (score GEMM + softmax epilogue) -> dV = P^T dO
            _lib.check(lib.adp_gemm_rows_softmax(q[b].data_ptr(), Dq, k[b].data_ptr(), T, T, 3, 0, scale, m[b].data_ptr(),
                                                 l[b].data_ptr(), None, None, P.data_ptr(), _sp()))
            tn(P, do[b], C, dvf, dv[b])
            # dS = scale * P * (dO V^T - delta) in the epilogue of the dP GEMM -> dQ = dS K, dK = dS^T Q
            _lib.check(lib.adp_gemm_rows_softmax(do[b].data_ptr(), C, v[b].data_ptr(), T, T, 4, 0, scale, None, None,
                                                 delta.data_ptr(), P.data_ptr(), D.data_ptr(), _sp()))
            _lib.check(lib.adp_gemm_rows_bf16(D.data_ptr(), T, None, 0, k[b].data_ptr(), 1, dq[b].data_ptr(), Dq, None, 0, None, T, _sp()))
            tn(D, q[b], Dq, dkf, dk[b])
        return dq, dk, dv, None


class _Residual(torch.autograd.Function):
    """feat + gamma * attended (:134)"""

    @staticmethod
    def forward(ctx, feat, att, gamma):
        C = feat.shape[-1]
        rows = feat.numel() // C
        y = torch.empty_like(feat)
        _lib.check(_lib.load().adp_rows_op(1, feat.data_ptr(), att.data_ptr(), gamma.data_ptr(), y.data_ptr(), rows, C, _sp()))
        ctx.save_for_backward(att, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        att, gamma = ctx.saved_tensors
        C = att.shape[-1]
        rows = att.numel() // C
        dy = dy.contiguous()
        lib = _lib.load()
        datt = torch.empty_like(att)
        _lib.check(lib.adp_rows_op(3, dy.data_ptr(), None, gamma.data_ptr(), datt.data_ptr(), rows, C, _sp()))
        dgamma = torch.empty(1, device=dy.device, dtype=torch.float32)
        _lib.check(lib.adp_rows_reduce(2, dy.data_ptr(), att.data_ptr(), rows, C, dgamma.data_ptr(), None, _sp()))
        return dy, datt, dgamma


class _ConvT2x2(torch.autograd.Function):
    """nn.ConvTranspose2d(C, N, kernel_size=2, stride=2) of the bilinear=False decoder (reference :65-66):
    y[b, 2i+a, 2j+c, n] = bias[n] + sum_k x[b, i, j, k] W[k][n][a][c].  Stride = kernel: the four taps never overlap, so
    the layer is ONE row GEMM with 4N columns (a, c, n) per input pixel on the tensor cores plus a pixel shuffle."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, H, W, C = x.shape
        N = weight.shape[1]
        rows = B * H * W
        lib = _lib.load()
        wcat = weight.detach().permute(0, 2, 3, 1).reshape(C, 4 * N).to(_BF16).contiguous()     # [K = C][N' = (a, c, n)]
        ys = torch.empty((rows, 4 * N), device=x.device, dtype=_BF16)
        _lib.check(lib.adp_gemm_rows_bf16(x.data_ptr(), C, None, 0, wcat.data_ptr(), 1, ys.data_ptr(), 4 * N, None, 0, None, rows, _sp()))
        y = torch.empty((B, 2 * H, 2 * W, N), device=x.device, dtype=_BF16)
        _lib.check(lib.adp_pixel_shuffle2(ys.data_ptr(), _ptr(bias), y.data_ptr(), B, H, W, N, 0, _sp()))
        ctx.save_for_backward(x, wcat)
        ctx.has_bias = bias is not None
        ctx.wshape = tuple(weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wcat = ctx.saved_tensors
        B, H, W, C = x.shape
        N = ctx.wshape[1]
        rows = B * H * W
        lib = _lib.load()
        dy = dy.contiguous()
        dys = torch.empty((rows, 4 * N), device=dy.device, dtype=_BF16)
        _lib.check(lib.adp_pixel_shuffle2(dy.data_ptr(), None, dys.data_ptr(), B, H, W, N, 1, _sp()))
        dx = torch.empty_like(x)          # dx[row][k] = sum_n' dys[row][n'] wcat[k][n']: wcat is the [N = C][K = 4N] "NT" operand
        _lib.check(lib.adp_gemm_rows_bf16(dys.data_ptr(), 4 * N, None, 0, wcat.data_ptr(), 0, dx.data_ptr(), C, None, 0, None, rows, _sp()))
        dwcat = torch.zeros((C, 4 * N), device=dy.device, dtype=torch.float32)
        _lib.check(lib.adp_gemm_tn_bf16(x.data_ptr(), C, dys.data_ptr(), 4 * N, dwcat.data_ptr(), 4 * N, rows, _sp()))
        dw = dwcat.view(C, 2, 2, N).permute(0, 3, 1, 2).contiguous()
        db = None
        if ctx.has_bias:
            dbp = torch.empty(4 * N, device=dy.device, dtype=torch.float32)
            ws = torch.empty(8 * N, device=dy.device, dtype=torch.float64)
            _lib.check(lib.adp_rows_reduce(0, dys.data_ptr(), None, rows, 4 * N, dbp.data_ptr(), ws.data_ptr(), _sp()))
            db = dbp.view(4, N).sum(0)
        return dx, dw, db


class _Interp(torch.autograd.Function):
    """F.interpolate(depth, size=(S, S), mode='bilinear', align_corners=False) (reference :322-328), fp32 planes."""

    @staticmethod
    def forward(ctx, y, size):
        B, C, H, W = y.shape
        y = y.contiguous()
        out = torch.empty((B, C, size, size), device=y.device, dtype=torch.float32)
        _lib.check(_lib.load().adp_bilinear_resize_forward(y.data_ptr(), out.data_ptr(), B * C, H, W, size, size, _sp()))
        ctx.shape = (B, C, H, W, size)
        return out

    @staticmethod
    def backward(ctx, dy):
        B, C, H, W, size = ctx.shape
        dy = dy.contiguous().float()
        dx = torch.empty((B, C, H, W), device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load().adp_bilinear_resize_backward(dy.data_ptr(), dx.data_ptr(), B * C, H, W, size, size, _sp()))
        return dx, None


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, max_depth):
        B, H, W, C = x.shape
        y = torch.empty((B, 1, H, W), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().adp_depth_head_forward(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), max_depth, B * H * W, C,
                                                      y.data_ptr(), _sp()))
        ctx.save_for_backward(x, weight, bias)
        ctx.max_depth = max_depth
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        B, H, W, C = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dw = torch.empty_like(weight)
        db = torch.empty_like(bias)
        _lib.check(_lib.load().adp_depth_head_backward(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), ctx.max_depth, dy.data_ptr(),
                                                       B * H * W, C, dx.data_ptr(), dw.data_ptr(), db.data_ptr(), _sp()))
        return dx, dw, db, None


# ============================================================================================ parameter containers
class DoubleConv(nn.Module):
    """(convolution => [BN] => ReLU) * 2   (reference :22-39)"""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))

    def run(self, x0, x1, training):
        seq = self.double_conv
        e = _Conv3x3.apply(x0, x1, seq[0].weight)
        a = _BnRelu.apply(e, seq[1].weight, seq[1].bias, None, seq[1], training)
        e = _Conv3x3.apply(a, None, seq[3].weight)
        return _BnRelu.apply(e, seq[4].weight, seq[4].bias, None, seq[4], training)

    def run_stem(self, x, channel, training):
        seq = self.double_conv
        e = _Conv3x3Stem.apply(x, channel, seq[0].weight)
        a = _BnRelu.apply(e, seq[1].weight, seq[1].bias, None, seq[1], training)
        e = _Conv3x3.apply(a, None, seq[3].weight)
        return _BnRelu.apply(e, seq[4].weight, seq[4].bias, None, seq[4], training)


class Down(nn.Module):
    """Downscaling with maxpool then double conv (reference :42-53)"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def run(self, x, training):
        return self.maxpool_conv[1].run(_MaxPool2.apply(x), None, training)


class Up(nn.Module):
    """Upscaling then double conv (reference :56-78)"""

    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels)

    def run(self, x1, x2, training):
        if isinstance(self.up, nn.ConvTranspose2d):
            x1 = _ConvT2x2.apply(x1, self.up.weight, self.up.bias)
        else:
            x1 = _Upsample2.apply(x1)
        if x1.shape[1:3] != x2.shape[1:3]:
            # (the F.pad of reference :72-73 only acts on sizes that are not multiples of 16; forward() admits powers of two)
            raise NotImplementedError("feature sizes that are not multiples of 16 (the F.pad of reference :72-73) are not built")
        return self.conv.run(x2, x1, training)          # torch.cat([x2, x1], dim=1) read as two tensors


class BinauralCrossAttention(nn.Module):
    """Cross-attention between left and right channel features (reference :81-153)"""

    def __init__(self, channels, reduction=8):
        super().__init__()
        self.channels = channels
        self.reduction = reduction
        self.query = nn.Conv2d(channels, channels // reduction, kernel_size=1)
        self.key = nn.Conv2d(channels, channels // reduction, kernel_size=1)
        self.value = nn.Conv2d(channels, channels, kernel_size=1)
        self.out = nn.Conv2d(channels, channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def _one_way(self, a, b):
        """a attends to b: a + gamma * out(softmax(q(a) k(b)^T / sqrt(C)) v(b))"""
        B, H, W, C = a.shape
        T = H * W
        q = _Conv1x1.apply(a, None, self.query.weight, self.query.bias).reshape(B, T, -1)
        k = _Conv1x1.apply(b, None, self.key.weight, self.key.bias).reshape(B, T, -1)
        v = _Conv1x1.apply(b, None, self.value.weight, self.value.bias).reshape(B, T, C)
        att = _Attend.apply(q, k, v, 1.0 / math.sqrt(C))
        att = _Conv1x1.apply(att.reshape(B, H, W, C), None, self.out.weight, self.out.bias)
        return _Residual.apply(a, att, self.gamma)

    def run(self, left, right):
        return self._one_way(left, right), self._one_way(right, left)


class BinauralEncoder(nn.Module):
    """Encoder for a single channel (reference :156-178)"""

    def __init__(self, base_channels=64, bilinear=True):
        super().__init__()
        self.inc = DoubleConv(1, base_channels)
        self.down1 = Down(base_channels, base_channels * 2)
        self.down2 = Down(base_channels * 2, base_channels * 4)
        self.down3 = Down(base_channels * 4, base_channels * 8)
        factor = 2 if bilinear else 1
        self.down4 = Down(base_channels * 8, base_channels * 16 // factor)

    def run(self, x, channel, training):
        x1 = self.inc.run_stem(x, channel, training)
        x2 = self.down1.run(x1, training)
        x3 = self.down2.run(x2, training)
        x4 = self.down3.run(x3, training)
        x5 = self.down4.run(x4, training)
        return {"x1": x1, "x2": x2, "x3": x3, "x4": x4, "x5": x5}


class BinauralAttentionDepthNet(nn.Module):
    """Binaural Attention Depth Estimation Network (reference :181-340)."""

    def __init__(self, base_channels=64, bilinear=True, output_size=256, max_depth=30.0, attention_levels=[2, 3, 4, 5]):
        super().__init__()
        if base_channels % 64:
            raise NotImplementedError("base_channels must be a multiple of 64 on the tcgen05 path")
        self.output_size = output_size
        self.max_depth = max_depth
        self.bilinear = bilinear
        self.attention_levels = attention_levels
        self.left_encoder = BinauralEncoder(base_channels, bilinear)
        self.right_encoder = BinauralEncoder(base_channels, bilinear)
        self.attention_modules = nn.ModuleDict()
        channel_map = {1: base_channels, 2: base_channels * 2, 3: base_channels * 4, 4: base_channels * 8,
                       5: base_channels * 8 if bilinear else base_channels * 16}
        for level in attention_levels:
            self.attention_modules[f"attn_{level}"] = BinauralCrossAttention(channels=channel_map[level], reduction=8)
        self.fusion_layers = nn.ModuleDict()
        for level in [1, 2, 3, 4, 5]:
            ch = channel_map[level]
            self.fusion_layers[f"fusion_{level}"] = nn.Sequential(nn.Conv2d(ch * 2, ch, kernel_size=1), nn.BatchNorm2d(ch),
                                                                  nn.ReLU(inplace=True))
        factor = 2 if bilinear else 1
        self.up1 = Up(base_channels * 16, base_channels * 8 // factor, bilinear)
        self.up2 = Up(base_channels * 8, base_channels * 4 // factor, bilinear)
        self.up3 = Up(base_channels * 4, base_channels * 2 // factor, bilinear)
        self.up4 = Up(base_channels * 2, base_channels, bilinear)
        self.outc = nn.Sequential(nn.Conv2d(base_channels, 1, kernel_size=1), nn.Sigmoid())
        self._init_weights()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                _channels_last_(m)

    def _init_weights(self):
        """reference :268-277"""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias, 0)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                _channels_last_(m)
        return out

    def load_state_dict(self, state_dict, strict=True, assign=False):
        if any(k.startswith("module.") for k in state_dict):
            state_dict = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                _channels_last_(m)
        return out

    def forward(self, x):
        """x: [B, 2, H, W] CUDA fp32 binaural spectrogram -> depth [B, 1, H, W] fp32 (reference :279-334)."""
        _lib.require_cuda(x, "BinauralAttentionDepthNet input", torch.float32)
        if x.dim() != 4 or x.shape[1] != 2 or x.shape[2] != x.shape[3]:
            raise ValueError("expected input [B,2,S,S], got %s" % (tuple(x.shape),))
        S = x.shape[-1]
        if S & (S - 1) or S < 128:
            raise NotImplementedError("input size must be a power of two >= 128 (every attention level needs >= 64 tokens)")
        x = x.contiguous()
        training = self.training
        left = self.left_encoder.run(x, 0, training)
        right = self.right_encoder.run(x, 1, training)
        fused = {}
        for level in [1, 2, 3, 4, 5]:
            lf, rf = left[f"x{level}"], right[f"x{level}"]
            if level in self.attention_levels:
                lf, rf = self.attention_modules[f"attn_{level}"].run(lf, rf)
            fl = self.fusion_layers[f"fusion_{level}"]
            e = _Conv1x1.apply(lf, rf, fl[0].weight, None)          # torch.cat([left, right], 1) read as two tensors
            fused[level] = _BnRelu.apply(e, fl[1].weight, fl[1].bias, fl[0].bias, fl[1], training)
        y = self.up1.run(fused[5], fused[4], training)
        y = self.up2.run(y, fused[3], training)
        y = self.up3.run(y, fused[2], training)
        y = self.up4.run(y, fused[1], training)
        depth = _Head.apply(y, self.outc[0].weight, self.outc[0].bias, float(self.max_depth))
        if depth.shape[-1] != self.output_size:
            depth = _Interp.apply(depth, int(self.output_size))
        # (torch.clamp(depth, 0, max_depth) of :331 is the identity here: sigmoid * max_depth lies inside the range and the
        # bilinear weights are a convex combination)
        return depth

    def get_num_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)
