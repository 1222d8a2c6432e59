"""GPU parity of the config-4 building blocks (models/binaural_attention_model.py) through the C ABI, against plain
PyTorch fp32 on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def lib():
    from audio_depth_estimation_b200 import _lib
    return _lib, _lib.load()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def nhwc_bf16(t):     # [B,C,H,W] fp32 -> bf16 NHWC storage, plus the rounded fp32 NCHW value
    q = t.to(torch.bfloat16)
    return q.permute(0, 2, 3, 1).contiguous(), q.float()


@pytest.mark.parametrize("B,H,W,C0,C1,N", [(2, 32, 32, 64, 0, 64), (1, 16, 16, 128, 128, 128), (3, 8, 8, 256, 0, 512),
                                           (2, 4, 4, 512, 512, 256), (1, 64, 64, 64, 64, 64), (2, 2, 2, 512, 0, 512)])
def test_conv3x3_fprop_dgrad_wgrad(B, H, W, C0, C1, N):
    _lib, L = lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + C0 + N)
    C = C0 + C1
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    w = torch.randn(N, C, 3, 3, device="cuda", generator=g) / (3 * C ** 0.5)
    dy = torch.randn(B, N, H, W, device="cuda", generator=g)
    xs, xr = nhwc_bf16(x)
    dys, dyr = nhwc_bf16(dy)
    w16 = w.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()          # [N][3][3][C]: channels_last memory of the weight
    wr = w.to(torch.bfloat16).float()
    x0 = xs[..., :C0].contiguous()
    x1 = xs[..., C0:].contiguous() if C1 else None
    scratch = torch.empty(B * H * W * max(N, C), device="cuda", dtype=torch.float32)
    sp = _lib.stream_ptr()
    # forward
    y = torch.empty(B, H, W, N, device="cuda", dtype=torch.bfloat16)
    tc0 = L.adp_tc_launch_count()
    _lib.check(L.adp_conv2d_k3s1_fprop(x0.data_ptr(), C0, x1.data_ptr() if C1 else None, C1, w16.data_ptr(), y.data_ptr(), B, H, W,
                                       N, scratch.data_ptr(), scratch.numel() * 4, sp))
    assert L.adp_tc_launch_count() > tc0
    ref = F.conv2d(xr, wr, padding=1)
    assert rel(y.float().permute(0, 3, 1, 2), ref) <= 6e-3
    # data gradient
    dx0 = torch.empty(B, H, W, C0, device="cuda", dtype=torch.bfloat16)
    dx1 = torch.empty(B, H, W, max(C1, 1), device="cuda", dtype=torch.bfloat16)
    _lib.check(L.adp_conv2d_k3s1_dgrad(dys.data_ptr(), N, w16.data_ptr(), dx0.data_ptr(), C0, dx1.data_ptr() if C1 else None, C1,
                                       B, H, W, scratch.data_ptr(), scratch.numel() * 4, sp))
    dref = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, padding=1)
    got = torch.cat([dx0, dx1], dim=-1) if C1 else dx0
    assert rel(got.float().permute(0, 3, 1, 2), dref) <= 6e-3
    # weight gradient (fp32, weight memory layout [N][3][3][C])
    dw = torch.full((N, 3, 3, C), 7.0, device="cuda", dtype=torch.float32)       # overwritten, not accumulated
    _lib.check(L.adp_conv2d_k3s1_wgrad(dys.data_ptr(), N, x0.data_ptr(), C0, x1.data_ptr() if C1 else None, C1, dw.data_ptr(), B,
                                       H, W, sp))
    wref = torch.nn.grad.conv2d_weight(xr, wr.shape, dyr, padding=1)
    assert rel(dw.permute(0, 3, 1, 2), wref) <= 2e-3


@pytest.mark.parametrize("M,K0,K1,N0,N1,b_kn,f32", [(256, 64, 0, 64, 0, 0, False), (1000, 128, 64, 128, 64, 0, False),
                                                    (4096, 64, 0, 4096, 0, 0, True), (300, 256, 0, 128, 0, 1, False),
                                                    (16, 512, 0, 512, 0, 1, True), (4096, 4096, 0, 128, 0, 1, False)])
def test_gemm_rows(M, K0, K1, N0, N1, b_kn, f32):
    _lib, L = lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(M + K0 + N0)
    K, N = K0 + K1, N0 + N1
    a = (torch.randn(M, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn((K, N) if b_kn else (N, K), device="cuda", generator=g).to(torch.bfloat16)
    a0 = a[:, :K0].contiguous()
    a1 = a[:, K0:].contiguous() if K1 else None
    ref = a.float() @ (b.float() if b_kn else b.float().t())
    sp = _lib.stream_ptr()
    if f32:
        c = torch.full((M, N), 3.0, device="cuda", dtype=torch.float32)
        _lib.check(L.adp_gemm_rows_bf16(a0.data_ptr(), K0, a1.data_ptr() if K1 else None, K1, b.data_ptr(), b_kn, None, N0, None, N1,
                                        c.data_ptr(), M, sp))
        assert rel(c, ref) <= 1e-4
    else:
        c0 = torch.zeros(M, N0, device="cuda", dtype=torch.bfloat16)
        c1 = torch.zeros(M, max(N1, 1), device="cuda", dtype=torch.bfloat16)
        _lib.check(L.adp_gemm_rows_bf16(a0.data_ptr(), K0, a1.data_ptr() if K1 else None, K1, b.data_ptr(), b_kn, c0.data_ptr(), N0,
                                        c1.data_ptr() if N1 else None, N1, None, M, sp))
        got = torch.cat([c0, c1], dim=1) if N1 else c0
        assert rel(got.float(), ref) <= 5e-3
