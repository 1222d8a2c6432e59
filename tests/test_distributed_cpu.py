"""World-size-2 gloo tests (CPU) of the data-parallel host logic: the staged gradient buckets of
GradientReducer cover every parameter exactly once and sum across ranks, and the global-batch loss is
recovered from the all-reduced statistics {N, sum|p-g|, sum d, sum d^2}."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_depth_estimation_b200 import synthetic
from oracle import loss_oracle as lo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cfg():
    return SimpleNamespace(dataset=SimpleNamespace(depth_norm=False, max_depth=30.0, images_size=128, preprocess="resize",
                                                   name="batvisionv2"),
                           mode=SimpleNamespace(criterion="Combined", l1_weight=0.237, silog_weight=0.637,
                                                silog_lambda=0.869, learning_rate=0.002),
                           model=SimpleNamespace(precision="fp32"))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from audio_depth_estimation_b200.models.unetbaseline_model import define_G
        from audio_depth_estimation_b200.training import GradientReducer, default_stage_groups
        torch.manual_seed(0)
        net = define_G(_cfg(), 2, 1, 16, "unet_128", "batch", False, gpu_ids=[])
        net._flatten(torch.device("cpu"))          # the flat layout itself is device independent
        flat_p, flat_g, slices = net.flat_buffers()
        red = GradientReducer(net, stages_per_group=2)
        assert red.world == world and net.grad_ready_hook is not None
        assert net.stage_groups == default_stage_groups(7, 2)
        # the stage slices tile the bulk region (hidden-layer convolution weights), the tail holds every small tensor
        tlo, thi = net.tail_slice()
        assert slices[0][0] == 0 and slices[-1][1] == tlo and thi == flat_p.numel()
        assert all(slices[i][1] == slices[i + 1][0] for i in range(len(slices) - 1))
        n_in_slices = sum(p.numel() for st in net.staged_parameters() for p in st)
        assert n_in_slices == sum(p.numel() for p in net.parameters())
        offs = dict(zip(map(id, net._flat["params"]), net._flat["offs"]))
        for q in net.parameters():
            assert (offs[id(q)] < tlo) == net._is_bulk(q)
        assert all((e - b) % 64 == 0 for b, e in slices)
        # parameters are views of the flat buffer, 4-D weights in channels_last ([Cout][kh][kw][Cin]) order
        w = net.levels()[1]["conv"].weight
        assert w.data_ptr() >= flat_p.data_ptr() and w.stride() == (16 * w.shape[1], 1, 4 * w.shape[1], w.shape[1])
        # broadcast makes replicas identical
        if rank == 1:
            flat_p.add_(1.0)
        red.broadcast_parameters(0)
        ref = [torch.zeros_like(flat_p) for _ in range(world)]
        dist.all_gather(ref, flat_p)
        assert torch.equal(ref[0], ref[1])
        # staged all-reduce == sum over ranks, each bucket launched as its stage group completes
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(flat_g.numel(), generator=g)
        flat_g.copy_(local)
        for gi in range(len(net.stage_groups)):
            net.grad_ready_hook(gi)
        red.wait()
        both = [torch.randn(flat_g.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        assert torch.allclose(flat_g, both[0] + both[1], atol=1e-6)
        # shard mode: reduce-scatter leaves the sum in the piece of every bucket this rank owns (and in the tail, which is
        # all-reduced); gathering the pieces restores the whole buffer on every rank
        sred = GradientReducer(net, stages_per_group=2, shard=True)
        assert sred.shard
        flat_g.copy_(local)
        for gi in range(len(net.stage_groups)):
            net.grad_ready_hook(gi)
        sred.wait()
        total = both[0] + both[1]
        for b, e in sred.owned_pieces() + [net.tail_slice()]:
            assert torch.allclose(flat_g[b:e], total[b:e], atol=1e-6)
        pieces = sred.owned_pieces()
        assert sum(e - b for b, e in pieces) * world == tlo
        sred.gather(flat_g)
        assert torch.allclose(flat_g, total, atol=1e-6)
        # global-batch loss from all-reduced statistics
        B = 2
        gt = synthetic.gt_depth(world * B, 64, 30.0, seed=9)
        rng = np.random.default_rng(10)
        pred = np.maximum(gt + rng.normal(0, 2.0, gt.shape), 0).astype(np.float32)
        sl = slice(rank * B, (rank + 1) * B)
        sums = torch.tensor(lo.loss_sums(pred[sl], gt[sl]), dtype=torch.float64)
        red.reduce_loss_sums(sums)
        loss, l1, si = lo.loss_from_sums(*sums.tolist(), 0.237, 0.637, 0.869)
        full, fl1, fsi, _ = lo.depth_loss_and_grad(pred, gt)
        assert abs(loss - full) <= 1e-12 * abs(full) and abs(l1 - fl1) <= 1e-12 and abs(si - fsi) <= 1e-12
        # ... which differs from the mean of per-rank losses (what naive DDP would compute)
        per_rank = lo.depth_loss_and_grad(pred[sl], gt[sl])[0]
        out.put((rank, loss, per_rank))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reducer_and_global_loss():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    res = sorted(out.get(timeout=10) for _ in range(world))
    assert abs(res[0][1] - res[1][1]) < 1e-12              # identical global loss on both ranks
    assert abs(0.5 * (res[0][2] + res[1][2]) - res[0][1]) > 1e-6   # != mean of per-rank losses


def test_stage_groups_cover_all_stages():
    from audio_depth_estimation_b200.training import default_stage_groups
    for nd in (7, 8):
        for k in (1, 2, 3, 16):
            groups = default_stage_groups(nd, k)
            assert groups[0][0] == 0 and groups[-1][1] == 2 * nd
            assert all(groups[i][1] == groups[i + 1][0] for i in range(len(groups) - 1))
