"""Two (or N) data-parallel ranks over NCCL against the CPU oracle evaluated with the reference's DataParallel semantics
(models/unetbaseline_model.py:52-55, train.py:642-669): batch sharded over the ranks, per-chunk BatchNorm statistics, the
loss formed on the GATHERED batch, gradients summed over the chunks, clip_grad_norm_ + AdamW on the global gradient.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/nccl_parity.py

Checks (fp32 mode: structure exact; bf16 mode: the stated 2e-2): identical loss on every rank and equal to the oracle's,
all-reduced gradients vs the oracle's per tensor, parameters bit-identical across ranks after the step (also with the
CUDA-graph step and the sharded optimiser), loss after three steps.  Imports oracle/: test infrastructure, run by
tests/test_gpu_nccl.py and by hand through gpurun --gpus 2 (log under profiles/).
"""
import os
import sys
from types import SimpleNamespace

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from audio_depth_estimation_b200 import synthetic  # noqa: E402
from audio_depth_estimation_b200.models.unetbaseline_model import define_G  # noqa: E402
from audio_depth_estimation_b200.training import TrainStep  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402


# Sigmoid head (cfg.dataset.depth_norm, the BatVision-V1 configuration): the Combined loss's own gradient is then well
# conditioned.  With the ReLU head dL/dy ~ 1/p on predictions within rounding of zero and gradients are only comparable
# to ~10 % between ANY two fp32 implementations (tests/test_gpu_parity.py).
DEPTH_NORM, MAX_DEPTH = True, 12.0


def cfg_for(precision, size):
    return SimpleNamespace(dataset=SimpleNamespace(depth_norm=DEPTH_NORM, max_depth=MAX_DEPTH, images_size=size, preprocess="resize",
                                                   name="batvisionv2"),
                           mode=SimpleNamespace(criterion="Combined", l1_weight=0.237, silog_weight=0.637, silog_lambda=0.869,
                                                learning_rate=0.002),
                           model=SimpleNamespace(precision=precision))


def oracle_dp_step(x, gt, sd, nd, world, lr, steps):
    """DataParallel semantics on the CPU: per-chunk forward (own batch statistics), loss on the gathered predictions."""
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    sdg = {k: v.clone() for k, v in sd.items()}
    names = [k for k in uo.ordered_state_dict(sdg, nd) if k.endswith((".weight", ".bias"))]
    for n in names:
        sdg[n].requires_grad_(True)
    m = [torch.zeros_like(sdg[n]) for n in names]
    v = [torch.zeros_like(sdg[n]) for n in names]
    xs, losses, grads0 = torch.from_numpy(x).chunk(world), [], None
    for it in range(1, steps + 1):
        for n in names:
            sdg[n].grad = None
        # (running statistics: every replica updates its own copy; rank 0's are what DataParallel keeps -- not compared)
        ys = [uo.unet_forward(xc, sdg, nd, DEPTH_NORM, training=True, update_running=False) for xc in xs]
        loss = uo.depth_loss(torch.cat(ys), torch.from_numpy(gt), depth_norm=DEPTH_NORM, max_depth=MAX_DEPTH)
        loss.backward()
        losses.append(float(loss.detach()))
        if grads0 is None:
            grads0 = {n: sdg[n].grad.detach().clone() for n in names}
        with torch.no_grad():
            uo.clip_adamw_step([sdg[n] for n in names], [sdg[n].grad for n in names], m, v, it, lr)
    return losses, grads0, {n: sdg[n].detach() for n in names}


def run_case(precision, graph, shard, rank, world, dev):
    nd, ngf, size, per_rank = 7, 16, 128, 2
    B = per_rank * world
    sd = uo.make_state_dict(ngf, nd, seed=77)
    x = synthetic.feature_like(B, size, seed=78)
    gt = synthetic.gt_depth(B, size, MAX_DEPTH, seed=79, normalised=DEPTH_NORM)
    steps = 3
    ref_losses, ref_g, ref_p = oracle_dp_step(x, gt, sd, nd, world, 0.002, steps)
    cfg = cfg_for(precision, size)
    net = define_G(cfg, 2, 1, ngf, "unet_128", "batch", False, gpu_ids=[dev.index])
    net.load_state_dict(uo.ordered_state_dict({k: v.clone() for k, v in sd.items()}, nd))
    step = TrainStep(cfg, net, lr=0.002, waveform_input=False, cuda_graph=graph, shard_optimizer=shard)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    xd, gd = torch.from_numpy(x[sl]).to(dev), torch.from_numpy(gt[sl]).to(dev)
    tol = 1e-3 if precision == "fp32" else 2e-2
    losses = []
    for it in range(steps):
        losses.append(float(step(xd, gd)))
        if it == 0 and not graph:
            # gradients of the first step: the all-reduced flat buffer holds the global-batch gradient on every rank
            worst = 0.0
            if not shard:
                for n, prm in net.named_parameters():
                    ref = ref_g[n if n in ref_g else "model." + n]
                    got = prm.grad.detach().cpu().double().reshape(ref.shape)
                    err = float((got - ref.double()).norm() / max(float(ref.double().norm()), 1e-30))
                    worst = max(worst, err)
                gtol = 2e-2 if precision == "fp32" else 0.35
                assert worst <= gtol, ("gradient mismatch", precision, worst)
            print("[rank %d] %s graph=%d shard=%d: worst per-tensor gradient error vs oracle %.3e" % (rank, precision, graph, shard, worst))
    torch.cuda.synchronize()
    step.sync_master_weights() if hasattr(step, "sync_master_weights") else None
    # identical loss on every rank (it is the GLOBAL-batch loss) and equal to the oracle's
    lt = torch.tensor(losses, device=dev, dtype=torch.float64)
    allv = [torch.zeros_like(lt) for _ in range(world)]
    dist.all_gather(allv, lt)
    for o in allv:
        assert torch.equal(o, allv[0]), ("losses differ across ranks", [a.tolist() for a in allv])
    assert abs(losses[0] - ref_losses[0]) <= tol * abs(ref_losses[0]), (precision, losses, ref_losses)
    ltol = 5e-3 if precision == "fp32" else 5e-2           # three AdamW steps: lr * sign(g) moves amplify rounding
    assert abs(losses[-1] - ref_losses[-1]) <= ltol * abs(ref_losses[-1]), (precision, losses, ref_losses)
    # parameters: bit-identical across ranks, and close to the oracle's after the steps
    flat_p = net.flat_buffers()[0]
    mine = flat_p.detach().clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), "parameters differ across ranks after %d steps" % steps
    worst_p = 0.0
    for n, prm in net.named_parameters():
        key = n if n in ref_p else "model." + n
        ref = ref_p[key]
        worst_p = max(worst_p, float((prm.detach().cpu() - ref).abs().max()))
    assert worst_p <= 3 * steps * 0.002, worst_p            # every element moves by at most ~lr per step
    print("[rank %d] %s graph=%d shard=%d: losses %s oracle %s, max |p - p_oracle| after %d steps %.2e" % (
        rank, precision, graph, shard, ["%.5f" % v for v in losses], ["%.5f" % v for v in ref_losses], steps, worst_p))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cases = [("fp32", False, False), ("bf16", False, False), ("fp32", True, False)]
    if os.environ.get("ADP_TEST_SHARD", "1") == "1":
        cases += [("fp32", False, True), ("bf16", True, True)]
    for precision, graph, shard in cases:
        run_case(precision, graph, shard, rank, world, dev)
        dist.barrier()
    if rank == 0:
        print("nccl parity ok: %d ranks" % world)
    sys.stdout.flush()
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)          # (NCCL communicators recorded into CUDA graphs make destroy_process_group() hang)


if __name__ == "__main__":
    main()
