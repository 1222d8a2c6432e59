"""Config 5 of BASELINE.json: UNetBaseline inference-only depth prediction sweep (eval-mode BatchNorm, waveform ->
feature -> depth), batch 1..1024, BatVision V1 and V2 shapes.  Prints one JSON line per (dataset, batch).
    python tools/bench_inference.py [--max-batch 1024] [--iters 20]"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from audio_depth_estimation_b200 import synthetic  # noqa: E402
from audio_depth_estimation_b200.feature import SpectrogramTransform  # noqa: E402
from audio_depth_estimation_b200.models.unetbaseline_model import define_G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    for name, L, dn, md in (("batvisionv2", synthetic.V2_LEN, False, 30.0), ("batvisionv1", synthetic.V1_LEN, True, 12.0)):
        cfg = SimpleNamespace(dataset=SimpleNamespace(name=name, depth_norm=dn, max_depth=md, images_size=256, preprocess="resize"),
                              model=SimpleNamespace(precision=args.precision))
        torch.manual_seed(0)
        net = define_G(cfg, 2, 1, 64, "unet_256", "batch", False, gpu_ids=[0]).eval()
        tr = SpectrogramTransform.for_cfg(cfg)
        B = 1
        while B <= args.max_batch:
            w = torch.from_numpy(synthetic.waveform(min(B, 64), L, seed=B)).to(dev)
            if B > 64:
                w = w.repeat((B + 63) // 64, 1, 1)[:B].contiguous()
            with torch.no_grad():
                for _ in range(3):
                    net(tr(w))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    y = net(tr(w))
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            print(json.dumps({"dataset": name, "batch": B, "ms_per_batch": ms, "samples_per_s": B / ms * 1e3,
                              "tflops": B * 11.93e9 / (ms * 1e-3) / 1e12, "dtype": args.precision}))
            sys.stdout.flush()
            B *= 2


if __name__ == "__main__":
    main()
