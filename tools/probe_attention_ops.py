"""Time the building blocks of one (sample, direction) of level-2 cross attention: T tokens, Dq = 64 (padded), C channels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from audio_depth_estimation_b200 import _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
lib = _lib.load()
sp = _lib.stream_ptr()
dev = "cuda"
q = torch.randn(T, 64, device=dev).to(torch.bfloat16)
k = torch.randn(T, 64, device=dev).to(torch.bfloat16)
v = torch.randn(T, C, device=dev).to(torch.bfloat16)
do = torch.randn(T, C, device=dev).to(torch.bfloat16)
S = torch.empty(T, T, device=dev, dtype=torch.float32)
P = torch.empty(T, T, device=dev, dtype=torch.bfloat16)
D = torch.empty(T, T, device=dev, dtype=torch.bfloat16)
m = torch.empty(T, device=dev, dtype=torch.int32)
mf = torch.empty(T, device=dev, dtype=torch.float32)
l = torch.empty(T, device=dev, dtype=torch.float32)
delta = torch.randn(T, device=dev)
o = torch.empty(T, C, device=dev, dtype=torch.bfloat16)
o64 = torch.empty(T, 64, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(T, C, device=dev, dtype=torch.float32)
scale = 0.05


def t(name, fn, flops=None, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%-34s %8.3f ms %s" % (name, ms, ("%7.1f TFLOP/s" % (flops / ms / 1e9)) if flops else ""), flush=True)


ck = _lib.check
ck(lib.adp_softmax_stats_init(m.data_ptr(), l.data_ptr(), T, sp))
sm = lambda mode, a, ka, b, by_col, out, pm=None: ck(lib.adp_gemm_rows_softmax(a.data_ptr(), ka, b.data_ptr(), T, T, mode, by_col, scale,
                                                                                   m.data_ptr(), l.data_ptr(), delta.data_ptr(),
                                                                                   pm.data_ptr() if pm is not None else None,
                                                                                   out.data_ptr() if out is not None else None, sp))
f_qk = 2.0 * T * T * 64
f_c = 2.0 * T * T * C
t("S fp32 = Q K^T (materialise)", lambda: ck(lib.adp_gemm_rows_bf16(q.data_ptr(), 64, None, 0, k.data_ptr(), 0, None, T, None, 0, S.data_ptr(), T, sp)), f_qk)
t("softmax_rows (old)", lambda: ck(lib.adp_softmax_rows(S.data_ptr(), T, T, scale, P.data_ptr(), mf.data_ptr(), l.data_ptr(), sp)))
t("fused mode 1 (row max)", lambda: sm(1, q, 64, k, 0, None), f_qk)
t("fused mode 2 (row sum)", lambda: sm(2, q, 64, k, 0, None), f_qk)
t("fused mode 3 (P, by row)", lambda: sm(3, q, 64, k, 0, P), f_qk)
t("fused mode 3 (P^T, by col)", lambda: sm(3, k, 64, q, 1, P), f_qk)
t("fused mode 4 (dS, by row)", lambda: sm(4, do, C, v, 0, D, P), f_c)
t("fused mode 4 (dS^T, by col)", lambda: sm(4, v, C, do, 1, D, P), f_c)
t("dP fp32 = dO V^T (materialise)", lambda: ck(lib.adp_gemm_rows_bf16(do.data_ptr(), C, None, 0, v.data_ptr(), 0, None, T, None, 0, S.data_ptr(), T, sp)), f_c)
t("softmax_backward (old)", lambda: ck(lib.adp_softmax_backward(P.data_ptr(), S.data_ptr(), T, T, scale, delta.data_ptr(), 0, D.data_ptr(), sp)))
t("O = P V (NN)", lambda: ck(lib.adp_gemm_rows_bf16(P.data_ptr(), T, None, 0, v.data_ptr(), 1, o.data_ptr(), C, None, 0, None, T, sp)), f_c)
t("dQ = dS K (NN, N=64)", lambda: ck(lib.adp_gemm_rows_bf16(D.data_ptr(), T, None, 0, k.data_ptr(), 1, o64.data_ptr(), 64, None, 0, None, T, sp)), f_qk)
t("dV = P^T dO (TN)", lambda: ck(lib.adp_gemm_tn_bf16(P.data_ptr(), T, do.data_ptr(), C, dw.data_ptr(), C, T, sp)), f_c)
dk = torch.zeros(T, 64, device=dev, dtype=torch.float32)
t("dK = dS^T Q (TN, N=64)", lambda: ck(lib.adp_gemm_tn_bf16(D.data_ptr(), T, q.data_ptr(), 64, dk.data_ptr(), 64, T, sp)), f_qk)
