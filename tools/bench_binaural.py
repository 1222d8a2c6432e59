"""BASELINE config 4: BinauralAttentionDepthNet forward+backward on one B200 (B in {2, 8}, 256x256 input,
attention_levels [3,4,5] and [2,3,4,5]).  Prints one JSON line per case (CUDA-event times, after warm-up).

    PYTHONPATH=. python tools/bench_binaural.py [--steps 5] [--warmup 2] [--size 256]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import json

import numpy as np
import torch

from audio_depth_estimation_b200 import _lib, synthetic
from audio_depth_estimation_b200.models.binaural_attention_model import BinauralAttentionDepthNet


def flops_per_sample(size, levels, base=64):
    """forward FLOPs (2*MAC) of the dense contractions: 3x3 convs, 1x1 convs, attention products."""
    ch = {1: base, 2: 2 * base, 3: 4 * base, 4: 8 * base, 5: 8 * base}
    hw = {l: (size >> (l - 1)) ** 2 for l in ch}
    f = 0.0
    enc = [(1, ch[1], 1), (ch[1], ch[1], 1), (ch[1], ch[2], 2), (ch[2], ch[2], 2), (ch[2], ch[3], 3), (ch[3], ch[3], 3),
           (ch[3], ch[4], 4), (ch[4], ch[4], 4), (ch[4], ch[5], 5), (ch[5], ch[5], 5)]
    f += 2 * sum(2.0 * 9 * a * b * hw[l] for a, b, l in enc)                       # two encoders
    dec = [(16 * base, 8 * base, 4 * base, 4), (8 * base, 4 * base, 2 * base, 3), (4 * base, 2 * base, base, 2), (2 * base, base, base, 1)]
    for cin, mid, cout, l in dec:
        f += 2.0 * 9 * cin * mid * hw[l] + 2.0 * 9 * mid * cout * hw[l]
    for l in ch:
        f += 2.0 * 2 * ch[l] * ch[l] * hw[l]                                          # fusion 1x1 (2C -> C)
    for l in levels:
        C, T = ch[l], hw[l]
        proj = 2.0 * C * (2 * (C // 8) + 2 * C) * T                                   # q, k, v, out
        att = 2.0 * T * T * (C // 8) + 2.0 * T * T * C                                # QK^T, PV
        f += 2 * (proj + att)                                                         # both directions
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--cases", default="3,4,5:2;3,4,5:8;2,3,4,5:2;2,3,4,5:8")
    a = ap.parse_args()
    lib = _lib.load()
    for case in a.cases.split(";"):
        lv, b = case.split(":")
        levels, B = [int(v) for v in lv.split(",")], int(b)
        torch.manual_seed(0)
        net = BinauralAttentionDepthNet(64, True, a.size, 30.0, levels).cuda().train()
        x = torch.from_numpy(synthetic.feature_like(B, a.size, seed=1)).cuda()
        r = torch.from_numpy(np.random.default_rng(2).normal(0, 1, (B, 1, a.size, a.size)).astype(np.float32)).cuda()
        for _ in range(a.warmup):
            (net(x) * r).sum().backward()
        torch.cuda.synchronize()
        l0, t0 = lib.adp_launch_count(), lib.adp_tc_launch_count()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        fwd = bwd = 0.0
        for _ in range(a.steps):
            for p in net.parameters():
                p.grad = None
            e0.record()
            y = net(x)
            e1.record()
            (y * r).sum().backward()
            e2.record()
            torch.cuda.synchronize()
            fwd += e0.elapsed_time(e1)
            bwd += e1.elapsed_time(e2)
        fwd /= a.steps
        bwd /= a.steps
        # the same step replayed from one CUDA graph (removes the Python / ctypes / tensor-map-encode time per launch)
        graph_ms = None
        for attempt in range(2):      # (the first capture of a process is invalidated by a one-time lazy initialisation)
            try:
                for p in net.parameters():
                    p.grad = None
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    (net(x) * r).sum().backward()
                gph.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.steps):
                    gph.replay()
                e1.record()
                torch.cuda.synchronize()
                graph_ms = e0.elapsed_time(e1) / a.steps
                break
            except Exception as exc:      # noqa: BLE001
                graph_ms = "failed: %s" % (str(exc)[:120],)
                torch.cuda.synchronize()
        fl = flops_per_sample(a.size, levels) * B
        print(json.dumps({"workload": "BinauralAttentionDepthNet fwd+bwd (config 4)", "attention_levels": levels, "batch": B, "size": a.size,
                          "params": net.get_num_params(), "fwd_ms": round(fwd, 2), "bwd_ms": round(bwd, 2),
                          "samples_per_s": round(B / (fwd + bwd) * 1e3, 2), "fwd_gflop": round(fl / 1e9, 1),
                          "fwd_tflops": round(fl / fwd / 1e9, 1), "fwd_bwd_tflops": round(3 * fl / (fwd + bwd) / 1e9, 1),
                          "graph_fwd_bwd_ms": round(graph_ms, 2) if isinstance(graph_ms, float) else graph_ms,
                          "graph_samples_per_s": round(B / graph_ms * 1e3, 2) if isinstance(graph_ms, float) else None,
                          "graph_tflops": round(3 * fl / graph_ms / 1e9, 1) if isinstance(graph_ms, float) else None,
                          "launches_per_step": (lib.adp_launch_count() - l0) // a.steps,
                          "tc_launches_per_step": (lib.adp_tc_launch_count() - t0) // a.steps,
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}), flush=True)
        del net, x, r, y
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
