"""Can a SWIZZLE_128B UMMA A operand start at an arbitrary 128-byte row of a larger shared-memory tile?
Prints, for every row shift 0..15 and descriptor base_offset in {0, shift & 7}, whether out == X[shift:shift+128] @ W^T."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_depth_estimation_b200 import _lib

lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(160, 64, device="cuda", generator=g).to(torch.bfloat16)
W = torch.randn(64, 64, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(128, 64, device="cuda", dtype=torch.float32)
for sbo in (1024,):
    for shift in range(0, 16):
        res = []
        for bo in sorted({0, shift & 7}):
            _lib.check(lib.adp_selftest_umma_offset(X.data_ptr(), W.data_ptr(), shift, bo, sbo, out.data_ptr(), _lib.stream_ptr()))
            torch.cuda.synchronize()
            ref = X[shift:shift + 128].float() @ W.float().t()
            err = float((out - ref).abs().max() / ref.abs().max())
            res.append("base_offset=%d: %s (%.2e)" % (bo, "OK " if err < 1e-2 else "BAD", err))
        print("sbo %4d shift %2d  " % (sbo, shift) + "   ".join(res), flush=True)

# non-power-of-two pitch between the 8-row groups (a halo tile whose image rows are 10 pixels wide): SBO = 1280 bytes
m = torch.arange(128, device="cuda")
for sbo, pitch in ((1280, 10), (1152, 9), (2048, 16)):
    for shift in (0, 1, 2):
        if shift + 15 * pitch + 8 > 160:
            continue
        _lib.check(lib.adp_selftest_umma_offset(X.data_ptr(), W.data_ptr(), shift, 0, sbo, out.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        rows = shift + (m // 8) * pitch + (m % 8)
        ref = X[rows].float() @ W.float().t()
        err = float((out - ref).abs().max() / ref.abs().max())
        print("sbo %4d (group pitch %2d rows) shift %d: %s (%.2e)" % (sbo, pitch, shift, "OK " if err < 1e-2 else "BAD", err), flush=True)
