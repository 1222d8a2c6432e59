#!/usr/bin/env python
"""bench.py -- train samples/sec of the hot path  waveform -> STFT log-magnitude feature ->
UNetBaseline fwd/bwd -> masked depth loss -> clip + AdamW  on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" is one full training step over one synthetic BatVision-V2-shaped batch (per-GPU batch 64:
BASELINE.json configs[1]; weak scaling, global batch 64*N, so N=8 is configs[2]'s global 512).
Prints ONE JSON line (rank 0).  `value`: inputs already resident in HBM; `e2e`: the same step through
the public API with pinned-host inputs copied H2D and the loss read back D2H every step.
`--impl reference` times the CPU oracle port of the reference path on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train samples/sec (STFT+UNet fwd/bwd+loss)"
UNIT = "samples/s"
FLOP_PER_SAMPLE_CONV = 35.38e9      # E2-E8, D8-D2 fwd+dgrad+wgrad


def make_cfg(precision="bf16"):
    return SimpleNamespace(
        dataset=SimpleNamespace(name="batvisionv2", depth_norm=False, max_depth=30.0, images_size=256,
                                preprocess="resize", audio_format="waveform"),
        mode=SimpleNamespace(criterion="Combined", l1_weight=0.237, silog_weight=0.637, silog_lambda=0.869,
                             learning_rate=0.002),
        model=SimpleNamespace(name="unet_baseline", generator="unet_256", precision=precision))


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step(B, threads):
    """One training step of the reference path on the CPU (oracle port): numpy STFT/log/min-max/resize,
    torch fp32 U-Net forward + loss + backward, clip + AdamW."""
    from audio_depth_estimation_b200 import synthetic
    from oracle import feature_oracle as fo
    from oracle import unet_oracle as uo
    torch.set_num_threads(threads)
    sd = uo.make_state_dict(64, 8, seed=0)
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))]
    for n in names:
        sd[n].requires_grad_(True)
    m = [torch.zeros_like(sd[n]) for n in names]
    v = [torch.zeros_like(sd[n]) for n in names]
    wave = synthetic.waveform(B, synthetic.V2_LEN, seed=1234)
    gt = torch.from_numpy(synthetic.gt_depth(B, 256, 30.0, seed=4321))
    state = {"step": 0}

    def step():
        state["step"] += 1
        x = torch.from_numpy(np.stack([fo.feature_v2(wave[b], 30.0, 256) for b in range(B)]))
        y = uo.unet_forward(x, sd, 8, False, training=True)
        loss = uo.depth_loss(y, gt)
        for n in names:
            sd[n].grad = None
        loss.backward()
        with torch.no_grad():
            uo.clip_adamw_step([sd[n] for n in names], [sd[n].grad for n in names], m, v, state["step"], 0.002)
        return float(loss.detach())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_batch
    step = cpu_reference_step(B, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    K = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    dt = (time.perf_counter() - t0) / K
    val = B / dt
    sample = "%d timed steps of batch %d (fp32, %d torch threads), oracle port of train.py:633-693" % (K, B, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UNetBaseline(unet_256, ngf 64, 54.4M params) full training step: STFT(512,64,16)+log+minmax+"
                               "resize -> U-Net fwd/bwd -> Combined L1+SIlog loss -> clip+AdamW; BatVision-V2 shapes "
                               "[B,2,7782] -> [B,1,256,256]",
                   "per_gpu_batch": 64, "cpu_sample_batch": B, "parallelism": "cpu"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    from audio_depth_estimation_b200 import _lib, synthetic
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    from audio_depth_estimation_b200.training import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B = args.batch
    cfg = make_cfg(args.precision)
    torch.manual_seed(0)
    net = define_G(cfg, 2, 1, 64, "unet_256", "batch", False, gpu_ids=[local])
    use_graph = not args.no_graph
    step = TrainStep(cfg, net, lr=0.002, stages_per_group=args.stages_per_group, cuda_graph=use_graph)
    wave_h = torch.from_numpy(synthetic.waveform(B, synthetic.V2_LEN, seed=1234 + rank)).pin_memory()
    gt_h = torch.from_numpy(synthetic.gt_depth(B, 256, 30.0, seed=4321 + rank)).pin_memory()
    wave_d, gt_d = wave_h.to(dev), gt_h.to(dev)
    step(wave_d, gt_d)                      # flattens parameters, allocates workspaces
    if world > 1:
        step.reducer.broadcast_parameters(0)
    launches0 = lib.adp_launch_count()
    step._eager_step(wave_d, gt_d)
    launches_per_step = lib.adp_launch_count() - launches0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident_step():
        step(wave_d, gt_d)

    # end to end: every step's inputs come from pinned host memory (copied on a side stream while the previous step
    # computes: DevicePrefetcher) and every step's loss is read back to the host
    import itertools
    from audio_depth_estimation_b200.training import DevicePrefetcher
    feed = DevicePrefetcher(itertools.repeat((wave_h, gt_h)), dev)

    def e2e_step():
        w, g = next(feed)
        return step(w, g).item()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()      # sampled from the warm-up on, through every timed region (all of it is under load)
    for _ in range(max(args.warmup, 3)):
        resident_step()
    if not use_graph:
        lib.adp_set_option(b"side_stream", 0)                  # the timed steps are also the profiled ones (see below)
        lib.adp_profile_enable(1)
    ms_total = timed(resident_step, args.steps)
    lib.adp_profile_enable(0)
    launches = launches_per_step * args.steps
    if use_graph:
        # the per-family CUDA events cannot live inside a replayed graph: the same K steps once more, eagerly
        prev_side = lib.adp_set_option(b"side_stream", 0)       # family event brackets must not overlap each other
        lib.adp_profile_enable(1)
        ms_eager = timed(lambda: step._eager_step(wave_d, gt_d), args.steps)
        lib.adp_profile_enable(0)
        lib.adp_set_option(b"side_stream", prev_side)
    else:
        ms_eager = ms_total
    pms, pwork, pcalls = (ctypes.c_double * 5)(), (ctypes.c_double * 5)(), (ctypes.c_longlong * 5)()
    _lib.check(lib.adp_profile_read(pms, pwork, pcalls))

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if sampler else None

    def finish():
        # NCCL communicators recorded into a CUDA graph make destroy_process_group() hang: drop the graph, make sure
        # every rank is done, then leave without running the collective teardown
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return
    pk = peaks()
    traffic = None
    tpath = os.path.join(REPO, "profiles", "r1_tc_dram_per_step.json")
    if os.path.exists(tpath) and B == 64 and args.precision == "bf16":
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]      # ncu capture of this workload, per launch
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    e2e = B * world / (ms_e2e / args.steps / 1e3)
    fam = ["gather_conv", "parity_convT", "wgrad", "thin", "elementwise"]
    families = {fam[k]: {"ms_per_step": pms[k] / args.steps, "tflops": (pwork[k] / (pms[k] * 1e-3) / 1e12) if pms[k] > 0 else None,
                         "calls_per_step": pcalls[k] / args.steps} for k in range(3)}
    conv_ms = sum(pms[k] for k in range(3))
    conv_flop = sum(pwork[k] for k in range(3))
    achieved = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "UNetBaseline(unet_256, ngf 64, 54.4M params) full training step: STFT(512,64,16)+log+minmax+"
                               "resize -> U-Net fwd/bwd -> Combined L1+SIlog loss -> clip+AdamW; BatVision-V2 shapes "
                               "[B,2,7782] -> [B,1,256,256]",
                   "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2_policy": "per-step working set (activations + 54.4M fp32 params/grads/moments, >1.5 GB) exceeds the 126 MB L2",
                   "tensor_core_path": bool(args.precision == "bf16"),
                   "launch": "whole step replayed from one CUDA graph" if use_graph else "eager launches",
                   "eager_ms_per_step": ms_eager / args.steps},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(wave_h.numel() * 4 + gt_h.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convolutions (gather + parity + wgrad families)",
                     "achieved": achieved, "peak": pk["tc_sustained"], "unit": "TFLOP/s",
                     "frac": (achieved / pk["tc_sustained"]) if achieved else None, "traffic": traffic,
                     "traffic_note": "mean DRAM bytes per tcgen05 conv launch (ncu, profiles/r1_tc_dram_per_step.json)",
                     "peak_source": pk["source"] + " (sustained bf16, kernel timed inside a long step)",
                     "share_of_step": conv_ms / ms_eager if ms_eager > 0 else None, "families": families,
                     "timed_over": "the K eager steps (per-family CUDA events on the launching stream)"},
        "step_tflops": value * 35.72e9 / 1e12,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = args.cpu_batch
        cstep = cpu_reference_step(cb, threads)
        cstep()
        t0 = time.perf_counter()
        cstep()
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cb / dt, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": "1 timed step of batch %d after 1 warm-up (fp32, %d torch threads)" % (cb, threads)}
    print(json.dumps(out))
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--stages-per-group", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
