"""Fused global-norm clip + AdamW for a UnetGenerator's flat parameter buffer.

One call = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) followed by
torch.optim.AdamW(model.parameters(), lr).step() (reference train.py:471-476, :689-691), executed
by two kernels over the flat fp32 parameter / gradient / moment buffers (adp_grad_sumsq,
adp_clip_adamw_step).  Defaults are torch.optim.AdamW's: betas (0.9, 0.999), eps 1e-8, weight
decay 0.01 on every parameter.
"""
import torch

from . import _lib

class _LrShim:
    """What torch.optim.lr_scheduler needs from an optimiser: `param_groups` whose 'lr' it reads and writes."""

    def _init_groups(self):
        self.param_groups = [dict(lr=self.lr, initial_lr=self.lr, betas=tuple(self.betas), eps=self.eps,
                                  weight_decay=self.weight_decay, params=[])]
        self.defaults = dict(lr=self.lr)
        self._opt_called = True

    def _pull_lr(self):
        self.lr = float(self.param_groups[0]["lr"])

    def set_lr(self, lr):
        self.lr = float(lr)
        self.param_groups[0]["lr"] = self.lr


class FusedClipAdamW(_LrShim):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0,
                 capturable=False, reducer=None):
        """capturable=True keeps the step counter on the device so that step() can be recorded in a CUDA graph.
        reducer: a training.GradientReducer in shard mode -- the gradients arrive reduce-scattered, this rank updates the
        1/world of every bucket it owns (plus the replicated tail of small tensors) and the new weights are all-gathered
        (bf16 mirror; fp32 masters in fp32 mode).  The other ranks' fp32 master shards and moments stay stale until
        sync_shards(), which state_dict() of the model and of the optimiser call."""
        self.model = model
        self.reducer = reducer
        self._stale = False
        if reducer is not None:
            model._master_sync = self.sync_shards
        self.capturable = bool(capturable)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.max_norm = float(max_norm) if max_norm is not None else 0.0
        self.step_count = 0
        self._state = None
        self.last_norm = None
        self._init_groups()

    def _buffers(self):
        p, g, _ = self.model.flat_buffers()
        st = self._state
        if st is None or st["p_ptr"] != p.data_ptr():
            old = st
            st = dict(p_ptr=p.data_ptr(), m=torch.zeros_like(p), v=torch.zeros_like(p),
                      sumsq=torch.zeros(1, device=p.device, dtype=torch.float64),
                      norm=torch.zeros(1, device=p.device, dtype=torch.float32),
                      step_dev=torch.full((1,), self.step_count, device=p.device, dtype=torch.int32),
                      scratch=torch.zeros(2, device=p.device, dtype=torch.float32))
            if old is not None and old["m"].numel() == p.numel():
                st["m"].copy_(old["m"])
                st["v"].copy_(old["v"])
            self._state = st
        # bf16 models: the update also writes bf16(p) into the model's mirror, which replaces the per-forward casts
        self._mirrored = getattr(self.model, "precision", None) == "bf16" and hasattr(self.model, "bf16_mirror")
        mirror = self.model.bf16_mirror() if self._mirrored else None

        def refs(ranges):
            arr = (_lib.TensorRef * len(ranges))()
            for i, (lo, hi) in enumerate(ranges):
                arr[i].p, arr[i].g = p.data_ptr() + 4 * lo, g.data_ptr() + 4 * lo
                arr[i].m, arr[i].v, arr[i].n = st["m"].data_ptr() + 4 * lo, st["v"].data_ptr() + 4 * lo, hi - lo
                arr[i].p_bf16 = (mirror.data_ptr() + 2 * lo) if mirror is not None else None
            return arr, len(ranges)
        return p, st, refs, mirror

    def zero_grad(self, set_to_none=True):
        pass  # gradients are overwritten by every backward pass

    def step(self, local_only=False):
        """local_only: (measurement) skip the cross-rank work of the sharded mode and just update what this rank owns."""
        lib = _lib.load()
        self._pull_lr()
        p, st, refs, mirror = self._buffers()
        self.step_count += 1
        red = self.reducer
        with torch.cuda.device(p.device):
            s = _lib.stream_ptr()
            st["sumsq"].zero_()
            if red is None:
                every, n = refs([(0, p.numel())])
                _lib.check(lib.adp_grad_sumsq(every, n, st["sumsq"].data_ptr(), s))
            else:
                # global norm: this rank's pieces summed over the ranks, plus the (replicated) tail once
                own, n_own = refs(red.owned_pieces())
                _lib.check(lib.adp_grad_sumsq(own, n_own, st["sumsq"].data_ptr(), s))
                if not local_only:
                    red.all_reduce_scalar(st["sumsq"])
                tail, n_tail = refs([self.model.tail_slice()])
                _lib.check(lib.adp_grad_sumsq(tail, n_tail, st["sumsq"].data_ptr(), s))
                every, n = refs(red.owned_pieces() + [self.model.tail_slice()])
            if self.capturable:
                _lib.check(lib.adp_clip_adamw_step_graph(every, n, st["sumsq"].data_ptr(), self.max_norm, self.lr,
                                                         self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                         st["step_dev"].data_ptr(), st["scratch"].data_ptr(),
                                                         st["norm"].data_ptr(), s))
            else:
                _lib.check(lib.adp_clip_adamw_step(every, n, st["sumsq"].data_ptr(), self.max_norm, self.lr,
                                                   self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                   self.step_count, st["norm"].data_ptr(), s))
            if red is not None and not local_only:
                red.gather(mirror if mirror is not None else p)        # the new weights, as the forward pass reads them
                self._stale = mirror is not None                       # (fp32 masters of the other ranks' pieces, m, v)
        self.model.mark_weights_dirty(by_optimizer=True)
        if self._mirrored:
            self.model.mirror_written()
        self.last_norm = st["norm"]
        return st["norm"]

    def sync_shards(self):
        """Sharded mode: all-gather the fp32 master weights and both AdamW moments (checkpoints, state_dict())."""
        if self.reducer is None or self._state is None:
            return
        p = self.model.flat_buffers()[0]
        if self._stale:
            self.reducer.gather(p)
            self._stale = False
        self.reducer.gather(self._state["m"])
        self.reducer.gather(self._state["v"])

    # ---- checkpointing: the layout torch.optim.AdamW.state_dict() has in the reference's checkpoints
    # (train.py:1006-1011 saves 'optimizer': optimizer.state_dict()), parameters in model.parameters() order.
    def _moment_views(self):
        _, st, _, _ = self._buffers()
        flat = self.model._flat
        off_of = {id(q): o for q, o in zip(flat["params"], flat["offs"])}
        return [(self.model._view_like(st["m"], off_of[id(q)], q), self.model._view_like(st["v"], off_of[id(q)], q))
                for q in self.model.parameters()]

    def state_dict(self):
        self.sync_shards()
        n = len(list(self.model.parameters()))
        state = {}
        if self._state is not None or self.model._flat is not None:
            if self.capturable and self._state is not None:
                self.step_count = int(self._state["step_dev"].item())
            if self.step_count > 0:
                for i, (m, v) in enumerate(self._moment_views()):
                    state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m.detach().clone().contiguous(),
                                "exp_avg_sq": v.detach().clone().contiguous()}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(n)))
        return {"state": state, "param_groups": [group], "max_norm": self.max_norm}

    def load_state_dict(self, sd):
        """Accepts the dictionary above, i.e. also a stock torch.optim.AdamW state_dict saved by the reference."""
        group = sd["param_groups"][0]
        self.set_lr(float(group.get("lr", self.lr)))
        self.betas = tuple(group.get("betas", self.betas))
        self.eps = float(group.get("eps", self.eps))
        self.weight_decay = float(group.get("weight_decay", self.weight_decay))
        if "max_norm" in sd:
            self.max_norm = float(sd["max_norm"])
        state = sd.get("state", {})
        if not state:
            self.step_count = 0
            return
        views = self._moment_views()
        if len(state) != len(views):
            raise ValueError("optimizer state has %d entries, the model %d parameters" % (len(state), len(views)))
        steps = set()
        with torch.no_grad():
            for i, (m, v) in enumerate(views):
                ent = state[i] if i in state else state[str(i)]
                m.copy_(ent["exp_avg"])
                v.copy_(ent["exp_avg_sq"])
                steps.add(int(float(ent["step"])))
        if len(steps) != 1:
            raise ValueError("per-parameter step counts differ: %s" % sorted(steps))
        self.step_count = steps.pop()
        self._state["step_dev"].fill_(self.step_count)


class FusedClipAdamWParams(_LrShim):
    """The same fused clip_grad_norm_(max_norm) + AdamW step for an arbitrary list of CUDA fp32 parameters (e.g. the
    config-4 BinauralAttentionDepthNet, whose parameters are ordinary separate tensors): two multi-tensor launches per
    24 tensors (adp_grad_sumsq, adp_clip_adamw_step).  Parameters must be dense (any memory format); a gradient whose
    memory order differs from its parameter's is re-laid-out first."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters")
        for p in self.params:
            _lib.require_cuda(p, "parameter", torch.float32)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.max_norm = float(max_norm) if max_norm is not None else 0.0
        self.step_count = 0
        self.exp_avg = [torch.zeros_like(p) for p in self.params]          # preserve_format: same memory order as p
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        dev = self.params[0].device
        self._sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self._norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.last_norm = None
        self._init_groups()

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def step(self):
        lib = _lib.load()
        self._pull_lr()
        live = [(p, m, v) for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq) if p.grad is not None]
        if not live:
            return None
        refs = (_lib.TensorRef * len(live))()
        keep = []
        for i, (p, m, v) in enumerate(live):
            g = p.grad
            # same element order in memory?  (1x1 kernels: channels_last and contiguous strides describe the same order)
            same = g.dtype == torch.float32 and (g.stride() == p.stride() or
                                                 (p.dim() == 4 and tuple(p.shape[2:]) == (1, 1) and g.is_contiguous()))
            if not same:
                g = torch.empty_like(p).copy_(g)
            keep.append(g)
            refs[i].p, refs[i].g, refs[i].m, refs[i].v, refs[i].n = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()
            refs[i].p_bf16 = None
        self.step_count += 1
        with torch.cuda.device(self.params[0].device):
            s = _lib.stream_ptr()
            self._sumsq.zero_()
            _lib.check(lib.adp_grad_sumsq(refs, len(live), self._sumsq.data_ptr(), s))
            _lib.check(lib.adp_clip_adamw_step(refs, len(live), self._sumsq.data_ptr(), self.max_norm, self.lr, self.betas[0],
                                               self.betas[1], self.eps, self.weight_decay, self.step_count, self._norm.data_ptr(), s))
        self.last_norm = self._norm
        return self._norm
