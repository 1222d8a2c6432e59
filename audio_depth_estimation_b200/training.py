"""The training / inference step of the reference loop on one GPU or on N data-parallel ranks.

`TrainStep` is the body of train.py:633-693 -- (GPU feature transform ->) model forward -> masked
criterion -> backward -> clip_grad_norm_(1.0) -> AdamW step -- with every numerical stage in
libadp_b200 and no host synchronisation inside the step.

Data parallelism replaces the reference's single-process nn.DataParallel
(models/unetbaseline_model.py:52-55: scatter, replicate 54 M parameters every forward, gather on
GPU 0) by one process per GPU over torch.distributed/NCCL:
  * batch-sharded inputs, replicated model, per-replica BatchNorm statistics (as DataParallel);
  * the four loss statistics {N, sum|p-g|, sum d, sum d^2} are all-reduced before the loss value and
    its gradient are formed, so the loss is the GLOBAL-batch loss the reference computes on the
    gathered predictions (train.py:642-669) -- not a mean of per-rank losses;
  * gradients are SUM-all-reduced per group of backward stages, launched while the later stages
    of backward are still running (NCCL's stream overlaps the compute stream);
  * the clip norm is computed after the all-reduce, hence identical on every rank.
"""
import os

import torch
import torch.distributed as dist

from .feature import SpectrogramTransform
from .optim import FusedClipAdamW, FusedClipAdamWParams
from .utils_criterion import METRIC_NAMES, batch_errors
from .utils_loss import DepthCriterion


# measurement only (tools/nccl_sweep.sh): ADP_SKIP_GRAD_ALLREDUCE=1 drops the gradient all-reduce (WRONG training) to
# separate the cost of the collective from the cost of running next to it
_SKIP_GRAD_ALLREDUCE = os.environ.get("ADP_SKIP_GRAD_ALLREDUCE", "0") == "1"


def default_stage_groups(num_downs, stages_per_group=2):
    n = 2 * num_downs
    return [(b, min(b + stages_per_group, n)) for b in range(0, n, stages_per_group)]


class GradientReducer:
    """Bucketed gradient reduction driven by UnetGenerator's backward stage groups.

    shard=False: SUM all-reduce of each stage group's slice of the flat gradient buffer (every rank then runs the whole
    optimiser step).  shard=True (ZeRO-1 style): SUM reduce-scatter of each slice -- rank r keeps the r-th 1/world of
    every bucket --, the optimiser updates only what the rank owns, and the updated weights come back with an
    all-gather of the bf16 mirror (half the bytes of the fp32 masters; fp32 mode gathers the masters).  The few-KB tail
    of small tensors is all-reduced and updated on every rank in both modes."""

    def __init__(self, model, process_group=None, stages_per_group=2, shard=False):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.shard = bool(shard) and self.world > 1 and 64 % self.world == 0      # (every slice is a multiple of 64 elements)
        self.pending = []
        model.stage_groups = default_stage_groups(model.num_downs, stages_per_group)
        model.grad_ready_hook = self._on_group_done if self.world > 1 else None

    def buckets(self):
        """[(lo, hi)] element ranges of the stage groups' slices (empty ones dropped)."""
        slices = self.model.flat_buffers()[2]
        out = []
        for b, e in self.model.stage_groups:
            lo, hi = slices[b][0], slices[e - 1][1]
            if hi > lo:
                out.append((lo, hi))
        return out

    def owned(self, lo, hi):
        piece = (hi - lo) // self.world
        return lo + self.rank * piece, lo + (self.rank + 1) * piece

    def owned_pieces(self):
        return [self.owned(lo, hi) for lo, hi in self.buckets()]

    def _on_group_done(self, gi):
        _, flat_g, slices = self.model.flat_buffers()
        b, e = self.model.stage_groups[gi]
        lo, hi = slices[b][0], slices[e - 1][1]
        if hi > lo and not _SKIP_GRAD_ALLREDUCE:
            if self.shard:
                olo, ohi = self.owned(lo, hi)
                self.pending.append(dist.reduce_scatter_tensor(flat_g[olo:ohi], flat_g[lo:hi], op=dist.ReduceOp.SUM,
                                                               group=self.group, async_op=True))
            else:
                self.pending.append(dist.all_reduce(flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        if gi == len(self.model.stage_groups) - 1 and not _SKIP_GRAD_ALLREDUCE:
            tlo, thi = self.model.tail_slice()
            if thi > tlo:
                self.pending.append(dist.all_reduce(flat_g[tlo:thi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_loss_sums(self, sums):
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_scalar(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def gather(self, flat):
        """All-gather every rank's owned piece of each bucket of `flat` (any flat buffer with the parameter layout)."""
        works = []
        for lo, hi in self.buckets():
            olo, ohi = self.owned(lo, hi)
            works.append(dist.all_gather_into_tensor(flat[lo:hi], flat[olo:ohi], group=self.group, async_op=True))
        for w in works:
            w.wait()

    def wait(self):
        for w in self.pending:
            w.wait()
        self.pending = []

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s weights and BatchNorm buffers."""
        if self.world <= 1:
            return
        flat_p, _, _ = self.model.flat_buffers()
        dist.broadcast(flat_p, src=src, group=self.group)
        for buf in self.model.buffers():
            dist.broadcast(buf, src=src, group=self.group)
        self.model.mark_weights_dirty()


class DevicePrefetcher:
    """Double-buffered host -> device input pipeline: while step i runs, the (pinned) host tensors of step i+1 are
    copied on a side stream.  Iterating yields device tensors that are safe to use on the current stream.

        for wave, gt in DevicePrefetcher(loader, device):       # loader yields tuples of pinned CPU tensors
            loss = step(wave, gt)
    """

    def __init__(self, batches, device):
        self.batches = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self._issue(0)

    def _issue(self, k):
        try:
            host = next(self.batches)
        except StopIteration:
            self.slots[k] = None
            return
        # the slot was last read by the step before the previous one, which the compute stream has passed
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            if self.slots[k] is None or any(d.shape != h.shape for d, h in zip(self.slots[k], host)):
                self.slots[k] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            for d, h in zip(self.slots[k], host):
                d.copy_(h, non_blocking=True)
            self.events[k].record(self.stream)

    def __iter__(self):
        return self

    def __next__(self):
        k = self.i & 1
        if self.slots[k] is None:
            raise StopIteration
        cur = self.slots[k]
        torch.cuda.current_stream(self.device).wait_event(self.events[k])
        self.i += 1
        self._issue(self.i & 1)          # next batch travels while the caller computes on `cur`
        return cur


class TrainStep:
    def __init__(self, cfg, model, lr=None, max_norm=1.0, process_group=None, stages_per_group=2,
                 waveform_input=True, cuda_graph=False, shard_optimizer=None):
        """cuda_graph=True records the whole step (feature -> forward -> loss -> backward -> clip+AdamW) once, after
        two eager warm-up steps, and replays it for every later batch of the same shape.  With several ranks the NCCL
        all-reduces are recorded into the graph as well (every rank must record and replay in lock-step)."""
        self.cfg = cfg
        self.cuda_graph = bool(cuda_graph)
        if shard_optimizer is None:
            # reduce-scatter + sharded clip/AdamW + bf16 all-gather pays from 4 ranks up (measured on 8 B200s: 4.19 against
            # 4.34 ms per step; a tie at 2 ranks); same weights either way (tools/nccl_parity.py)
            shard_optimizer = dist.is_initialized() and dist.get_world_size(process_group) >= 4
        self.shard_optimizer = bool(shard_optimizer)
        self._graph = None
        self._static = None
        self._eager_calls = 0
        self._epoch = None
        self._graph_lr = None
        self.model = model
        self.transform = SpectrogramTransform.for_cfg(cfg) if waveform_input else None
        self.reducer = GradientReducer(model, process_group, stages_per_group, shard=self.shard_optimizer)
        self.shard_optimizer = self.reducer.shard
        self.criterion = DepthCriterion.from_cfg(cfg, reduce_fn=self.reducer.reduce_loss_sums
                                                 if self.reducer.world > 1 else None)
        lr = lr if lr is not None else getattr(cfg.mode, "learning_rate", 1e-3)
        opt_name = str(getattr(cfg.mode, "optimizer", "AdamW"))
        if opt_name != "AdamW":             # train.py:471-476 also offers Adam / SGD; only the default is fused here
            raise NotImplementedError("cfg.mode.optimizer=%r: only 'AdamW' (the reference default) is implemented" % opt_name)
        self.optimizer = FusedClipAdamW(model, lr=lr, max_norm=max_norm, capturable=self.cuda_graph,
                                        reducer=self.reducer if self.reducer.shard else None)

    def features(self, batch):
        return self.transform(batch) if self.transform is not None else batch

    def sync_master_weights(self):
        """Sharded optimiser: all-gather the fp32 master weights (and the AdamW moments) every rank updated for the others.
        Called before state_dict() / checkpoints automatically; a no-op otherwise."""
        self.optimizer.sync_shards()

    def __call__(self, batch, gtdepth):
        """batch: waveform [B,2,L] (or features [B,2,S,S] when waveform_input=False), gtdepth [B,1,S,S];
        CUDA fp32.  Returns the loss as a 0-dim device tensor (no synchronisation)."""
        if self.cuda_graph:
            return self._graphed_step(batch, gtdepth)
        return self._eager_step(batch, gtdepth)

    def _eager_step(self, batch, gtdepth):
        self.model.train()
        x = self.features(batch)
        pred = self.model(x)
        loss = self.criterion(pred, gtdepth)
        loss.backward()
        self.reducer.wait()
        self.optimizer.step()
        return loss.detach()

    def _graphed_step(self, batch, gtdepth):
        key = (tuple(batch.shape), tuple(gtdepth.shape), batch.device)
        if self._static is not None and (self._static[0] != key or self._epoch != self.model._external_epoch
                                         or self._graph_lr != self.optimizer.lr):
            # new shape, a new learning rate (set_lr / a scheduler: lr is a captured kernel argument), or the parameters
            # were replaced behind the graph's back (load_state_dict, broadcast): record again
            self._graph, self._static, self._eager_calls = None, None, 0
        if self._graph is None:
            if self._eager_calls < 2:                  # workspaces, tensor maps, func attributes: all set up eagerly
                self._eager_calls += 1
                return self._eager_step(batch, gtdepth)
            # (no side-stream warm-up step here: it would be an extra, real optimiser update)
            sb, sg = batch.clone(), gtdepth.clone()
            torch.cuda.synchronize(batch.device)
            graph = torch.cuda.CUDAGraph()
            count = self.optimizer.step_count
            with torch.cuda.graph(graph):
                sloss = self._eager_step(sb, sg)
            self.optimizer.step_count = count          # recording is not a step (the device-side counter did not move)
            self._graph, self._static = graph, (key, sb, sg, sloss)
            self._graph_lr = self.optimizer.lr
            self._epoch = self.model._external_epoch
        _, sb, sg, sloss = self._static
        sb.copy_(batch, non_blocking=True)
        sg.copy_(gtdepth, non_blocking=True)
        self._graph.replay()
        self.optimizer.step_count += 1
        return sloss.clone()           # (the static tensor is overwritten by the next replay)

    @torch.no_grad()
    def evaluate(self, batch, gtdepth, metrics=False, protocol="train"):
        """test.py:231-241 / train.py:770-838: eval-mode forward + masked criterion; metrics=True adds the [B,7]
        per-sample error table (utils_criterion.batch_errors; still no host synchronisation)."""
        self.model.eval()
        pred = self.model(self.features(batch))
        loss = self.criterion(pred, gtdepth)
        if metrics:
            return pred, loss, batch_errors(gtdepth, pred, self.cfg, protocol=protocol)
        return pred, loss


@torch.no_grad()
def evaluate_loader(step, batches, protocol="test"):
    """The evaluation loop of test.py:231-300 (or the validation pass of train.py:770-845 with protocol='train')
    over an iterable of device (input, gt) pairs.  One device -> host copy at the end instead of two per sample.
    Returns {'loss': mean batch loss, 'abs_rel', 'rmse', 'delta1', 'delta2', 'delta3', 'log10', 'mae': sample means}."""
    losses, tables = [], []
    for batch, gt in batches:
        _, loss, errs = step.evaluate(batch, gt, metrics=True, protocol=protocol)
        losses.append(loss.reshape(1).double())
        tables.append(errs)
    if not tables:
        raise ValueError("evaluate_loader: no batches")
    row = torch.cat([torch.cat(losses).mean().reshape(1), torch.cat(tables).mean(0)]).tolist()
    return dict(zip(("loss",) + METRIC_NAMES, row))


class ModuleTrainStep:
    """The step body of train_binaural_attention.py:430-438 (zero_grad -> forward -> masked criterion -> backward -> AdamW;
    that trainer does NOT clip gradients, so max_norm defaults to None) for a model whose parameters are ordinary separate
    tensors, e.g. the config-4 BinauralAttentionDepthNet: `DepthCriterion` + `FusedClipAdamWParams`.  `features`
    (optional) is applied to the batch first, e.g. `SpectrogramTransform.for_cfg(cfg)`.  A torch LR scheduler can drive
    the step through `optimizer.param_groups` (CosineAnnealingLR / StepLR as in :300-312) or `optimizer.set_lr()`.
    Single process; returns the loss as a 0-dim device tensor."""

    def __init__(self, cfg, model, lr=None, max_norm=None, features=None, weight_decay=1e-2):
        self.cfg = cfg
        self.model = model
        self.features = features
        self.criterion = DepthCriterion.from_cfg(cfg)
        lr = lr if lr is not None else getattr(cfg.mode, "learning_rate", 1e-3)
        self.optimizer = FusedClipAdamWParams(model.parameters(), lr=lr, max_norm=max_norm, weight_decay=weight_decay)

    def __call__(self, batch, gtdepth):
        self.model.train()
        x = self.features(batch) if self.features is not None else batch
        self.optimizer.zero_grad()
        loss = self.criterion(self.model(x), gtdepth)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    @torch.no_grad()
    def evaluate(self, batch, gtdepth, metrics=False, protocol="train"):
        self.model.eval()
        pred = self.model(self.features(batch) if self.features is not None else batch)
        loss = self.criterion(pred, gtdepth)
        if metrics:
            return pred, loss, batch_errors(gtdepth, pred, self.cfg, protocol=protocol)
        return pred, loss
