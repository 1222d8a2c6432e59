"""Fused global-norm clip + AdamW for a UnetGenerator's flat parameter buffer.

One call = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) followed by
torch.optim.AdamW(model.parameters(), lr).step() (reference train.py:471-476, :689-691), executed
by two kernels over the flat fp32 parameter / gradient / moment buffers (adp_grad_sumsq,
adp_clip_adamw_step).  Defaults are torch.optim.AdamW's: betas (0.9, 0.999), eps 1e-8, weight
decay 0.01 on every parameter.
"""
import torch

from . import _lib


class FusedClipAdamW:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0,
                 capturable=False):
        """capturable=True keeps the step counter on the device so that step() can be recorded in a CUDA graph."""
        self.model = model
        self.capturable = bool(capturable)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.max_norm = float(max_norm) if max_norm is not None else 0.0
        self.step_count = 0
        self._state = None
        self.last_norm = None

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
        ref = (_lib.TensorRef * 1)()
        ref[0].p, ref[0].g, ref[0].m, ref[0].v, ref[0].n = (p.data_ptr(), g.data_ptr(), st["m"].data_ptr(),
                                                           st["v"].data_ptr(), p.numel())
        return p, st, ref

    def zero_grad(self, set_to_none=True):
        pass  # gradients are overwritten by every backward pass

    def step(self):
        lib = _lib.load()
        p, st, ref = self._buffers()
        self.step_count += 1
        with torch.cuda.device(p.device):
            s = _lib.stream_ptr()
            st["sumsq"].zero_()
            _lib.check(lib.adp_grad_sumsq(ref, 1, st["sumsq"].data_ptr(), s))
            if self.capturable:
                _lib.check(lib.adp_clip_adamw_step_graph(ref, 1, st["sumsq"].data_ptr(), self.max_norm, self.lr,
                                                         self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                         st["step_dev"].data_ptr(), st["scratch"].data_ptr(),
                                                         st["norm"].data_ptr(), s))
            else:
                _lib.check(lib.adp_clip_adamw_step(ref, 1, st["sumsq"].data_ptr(), self.max_norm, self.lr,
                                                   self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                   self.step_count, st["norm"].data_ptr(), s))
        self.model.mark_weights_dirty()
        self.last_norm = st["norm"]
        return st["norm"]

    def state_dict(self):
        st = self._state
        return dict(step=self.step_count, lr=self.lr, betas=self.betas, eps=self.eps,
                    weight_decay=self.weight_decay, max_norm=self.max_norm,
                    exp_avg=None if st is None else st["m"].clone(),
                    exp_avg_sq=None if st is None else st["v"].clone())

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.lr = float(sd.get("lr", self.lr))
        if sd.get("exp_avg") is not None:
            _, st, _ = self._buffers()
            st["m"].copy_(sd["exp_avg"])
            st["v"].copy_(sd["exp_avg_sq"])
