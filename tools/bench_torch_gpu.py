"""Comparator (NOT on the product path): the reference's own `define_G(unet_256)` under stock PyTorch + cuDNN on the same
B200, same batch and step body as bench.py -- the bar SURVEY.md 2.1 names ("the unmodified reference modules running
under stock torch + cuDNN on one B200, fp32 and bf16 autocast").

    python tools/bench_torch_gpu.py [--batch 64] [--steps 20] [--warmup 5] [--out profiles/r2_torch_cudnn_b200.json]

Step body (train.py:633-693): features (torchaudio Spectrogram(512, 64, 16) -> log -> per-channel min-max ->
antialiased bilinear resize, batched on the GPU -- already kinder than the reference's per-sample CPU loop), U-Net
forward, masked Combined loss, backward, clip_grad_norm_(1.0), AdamW.  Variants: fp32 (TF32 off / on), bf16 autocast
with channels_last, each eager and with the fused AdamW; optionally the whole step under torch.compile is NOT measured
(the reference does not use it).  Prints one JSON line per variant and writes them to --out.
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from audio_depth_estimation_b200 import synthetic  # noqa: E402
from oracle import reference_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r2_torch_cudnn_b200.json"))
    args = ap.parse_args()
    root = reference_step.reference_root()
    if root is None:
        raise SystemExit("reference modules not found (run oracle/vendor_reference.py in the build container)")
    define_G, SIlogLoss, _, _ = reference_step._import_reference(root)
    import torchaudio.transforms as T
    dev = torch.device("cuda", 0)
    B = args.batch
    wave = torch.from_numpy(synthetic.waveform(B, synthetic.V2_LEN, seed=1234)).to(dev)
    gt = torch.from_numpy(synthetic.gt_depth(B, 256, 30.0, seed=4321)).to(dev)
    cfg = SimpleNamespace(dataset=SimpleNamespace(depth_norm=False, preprocess="resize", images_size=256, max_depth=30.0))
    spec_t = T.Spectrogram(n_fft=512, win_length=64, power=1.0, hop_length=16).to(dev)

    def features():
        s = torch.log(spec_t(wave) + 1e-8)                                    # [B,2,257,487]
        lo, hi = s.amin(dim=(2, 3), keepdim=True), s.amax(dim=(2, 3), keepdim=True)
        s = torch.where(hi > lo, (s - lo) / (hi - lo), torch.zeros_like(s))
        return F.interpolate(s, size=(256, 256), mode="bilinear", antialias=True, align_corners=False)

    results = []
    for name, amp, tf32, channels_last in (("fp32 (TF32 off)", False, False, False), ("fp32 (TF32 on)", False, True, False),
                                           ("bf16 autocast, channels_last", True, True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        torch.manual_seed(0)
        net = define_G(cfg, 2, 1, 64, "unet_256", "batch", False, gpu_ids=[]).to(dev)
        if channels_last:
            net = net.to(memory_format=torch.channels_last)
        l1c, sic = torch.nn.L1Loss(), SIlogLoss(lambda_scale=0.869)
        opt = torch.optim.AdamW(net.parameters(), lr=0.002, fused=True)

        def step():
            net.train()
            x = features()
            if channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                y = net(x)
            y = y.float()
            mask = gt != 0.0
            loss = 0.237 * l1c(y[mask], gt[mask]) + 0.637 * sic(y[mask], gt[mask])
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
            opt.step()
            return loss

        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        # the same step replayed from a CUDA graph (what bench.py's `value` does for the hand-written path)
        graph_ms = None
        try:
            g = torch.cuda.CUDAGraph()
            sopt = torch.optim.AdamW(net.parameters(), lr=0.002, capturable=True, fused=True)

            def gstep():
                x = features()
                if channels_last:
                    x = x.contiguous(memory_format=torch.channels_last)
                sopt.zero_grad(set_to_none=False)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    y = net(x)
                y = y.float()
                w = (gt != 0.0).float()                           # (boolean-mask gathers have data-dependent shapes: not capturable)
                n = w.sum()
                p, q = y.clamp(min=1e-6), gt.clamp(min=1e-6)
                d = (torch.log(p) - torch.log(q)) * w
                m1, m2 = d.sum() / n, (d * d).sum() / n
                loss = 0.237 * ((y - gt).abs() * w).sum() / n + 0.637 * torch.sqrt((m2 - 0.869 * m1 * m1).clamp(min=0))
                loss.backward()
                torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
                sopt.step()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    gstep()
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                gstep()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_ms = e0.elapsed_time(e1) / args.steps
        except Exception as exc:                                  # report, do not hide
            graph_ms = "not capturable: %s" % str(exc).split("\n")[0][:160]
        row = {"variant": "reference define_G(unet_256) under torch %s + cuDNN %s: %s" % (torch.__version__, torch.backends.cudnn.version(), name),
               "batch": B, "eager_ms_per_step": ms, "eager_samples_per_s": B / ms * 1e3,
               "cuda_graph_ms_per_step": graph_ms,
               "cuda_graph_samples_per_s": (B / graph_ms * 1e3) if isinstance(graph_ms, float) else None,
               "gpu": torch.cuda.get_device_name(0), "steps": args.steps}
        print(json.dumps(row))
        results.append(row)
        del net, opt
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
