"""Time the 3x3 convolution family of config 4 per layer shape (B = 8, 256 x 256 input)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_depth_estimation_b200 import _lib

lib = _lib.load()
sp = _lib.stream_ptr()
B = 8
shapes = [(256, 64, 0, 64), (128, 64, 0, 128), (128, 128, 0, 128), (64, 128, 0, 256), (64, 256, 0, 256), (32, 256, 0, 512),
          (32, 512, 0, 512), (16, 512, 0, 512), (32, 512, 512, 512), (64, 256, 256, 256), (128, 128, 128, 128), (256, 64, 64, 64)]


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for H, C0, C1, N in shapes:
    C = C0 + C1
    x0 = torch.randn(B, H, H, C0, device="cuda").to(torch.bfloat16)
    x1 = torch.randn(B, H, H, C1, device="cuda").to(torch.bfloat16) if C1 else None
    w = torch.randn(N, 3, 3, C, device="cuda").to(torch.bfloat16)
    y = torch.empty(B, H, H, N, device="cuda", dtype=torch.bfloat16)
    dx0, dx1 = torch.empty_like(x0), (torch.empty_like(x1) if C1 else None)
    dw = torch.empty(N, 3, 3, C, device="cuda")
    p = lambda a: a.data_ptr() if a is not None else None
    fl = 2.0 * B * H * H * 9 * C * N
    f = t(lambda: _lib.check(lib.adp_conv2d_k3s1_fprop(p(x0), C0, p(x1), C1, p(w), p(y), B, H, H, N, None, 0, sp)))
    d = t(lambda: _lib.check(lib.adp_conv2d_k3s1_dgrad(p(y), N, p(w), p(dx0), C0, p(dx1), C1, B, H, H, None, 0, sp)))
    g = t(lambda: _lib.check(lib.adp_conv2d_k3s1_wgrad(p(y), N, p(x0), C0, p(x1), C1, p(dw), B, H, H, sp)))
    print("%3dx%-3d %4d+%-4d -> %-4d  fprop %7.3f ms %6.0f TF | dgrad %7.3f ms %6.0f TF | wgrad %7.3f ms %6.0f TF" %
          (H, H, C0, C1, N, f, fl / f / 1e9, d, fl / d / 1e9, g, fl / g / 1e9), flush=True)
