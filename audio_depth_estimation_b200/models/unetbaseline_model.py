"""B200 mirror of models/unetbaseline_model.py (reference :1-235): define_G / UnetGenerator /
UnetSkipConnectionBlock with the same constructor signatures, the same module tree and therefore
the same state_dict keys and shapes -- but `UnetGenerator.forward` runs the
whole encoder-decoder through libadp_b200 (adp_unet_forward / adp_unet_backward_stages) instead of
dispatching 16 cuDNN convolutions, 13 batch-norms, 7 torch.cat and the in-place activations.

The nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d objects are parameter containers only.  On
the first CUDA forward every parameter is re-homed into one flat fp32 buffer (4-D weights as
channels_last views, i.e. physically [Cout][kh][kw][Cin] / [Cin][kh][kw][Cout]), ordered by the
backward stage that finalises its gradient, so that gradient all-reduce buckets and the fused
clip+AdamW step are contiguous slices.  There is no CPU / eager fallback.
"""
import ctypes
import functools
import os

import torch
import torch.nn as nn
from torch.nn import init

from .. import _lib

_ALIGN = 64  # elements (256 B): every tensor starts on a 256-byte boundary inside the flat buffers


# ----------------------------------------------------------------------------- helpers (reference :9-120)
def init_weights(net, init_type="normal", init_gain=0.02):
    """Weight initialisation of reference :9-40: conv/linear weights by `init_type`,
    BatchNorm weight ~ N(1, gain), biases 0."""
    def visit(m):
        name = m.__class__.__name__
        if hasattr(m, "weight") and m.weight is not None and ("Conv" in name or "Linear" in name):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif "BatchNorm2d" in name:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)
    print("initialize network with %s" % init_type)
    net.apply(visit)
    if isinstance(net, UnetGenerator):
        net.mark_weights_dirty()


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):
    """Reference :42-57 moves the net to gpu_ids[0] and wraps it in nn.DataParallel.  Here the
    net moves to the device and multi-GPU is one process per GPU (data_parallel.DataParallelTrainer);
    checkpoints written by the reference ('module.' prefix) still load (see load_state_dict)."""
    if len(gpu_ids) > 0:
        if not torch.cuda.is_available():
            raise AssertionError("gpu_ids given but CUDA is not available")
        net.to(torch.device("cuda", gpu_ids[0]))
    init_weights(net, init_type, init_gain=init_gain)
    return net


def get_norm_layer(norm_type="instance"):
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    if norm_type in ("instance", "none"):
        raise NotImplementedError("normalization layer [%s]: only 'batch' (the UNetBaseline configuration, "
                                  "train.py:381) is implemented on the B200 path" % norm_type)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def define_G(cfg, input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal",
             init_gain=0.02, gpu_ids=[]):
    """Reference :84-120."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netG == "unet_128":
        net = UnetGenerator(cfg, input_nc, output_nc, 7, ngf, norm_layer=norm_layer, use_dropout=use_dropout)
    elif netG == "unet_256":
        net = UnetGenerator(cfg, input_nc, output_nc, 8, ngf, norm_layer=norm_layer, use_dropout=use_dropout)
    else:
        raise NotImplementedError("Generator model name [%s] is not recognized" % netG)
    return init_net(net, init_type, init_gain, gpu_ids)


# ----------------------------------------------------------------------------- blocks
class UnetSkipConnectionBlock(nn.Module):
    """Parameter container with the reference's Sequential layout (reference :157-235):
    outermost [Conv, sub, ReLU, ConvT(bias), ReLU|Sigmoid]; middle [LeakyReLU, Conv, BN, sub, ReLU,
    ConvT, BN]; innermost [LeakyReLU, Conv, ReLU, ConvT, BN]."""

    def __init__(self, cfg, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        if use_dropout:
            raise NotImplementedError("use_dropout=True is never used by the reference trainers (train.py:382)")
        norm_cls = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        if norm_cls is not nn.BatchNorm2d:
            raise NotImplementedError("only BatchNorm2d blocks are implemented on the B200 path")
        self.cfg = cfg
        self.outermost, self.innermost = outermost, innermost
        self.outer_nc, self.inner_nc = outer_nc, inner_nc
        self.input_nc = outer_nc if input_nc is None else input_nc
        conv = nn.Conv2d(self.input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=False)
        if outermost:
            up = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            tail = nn.Sigmoid() if cfg.dataset.depth_norm else nn.ReLU()
            layers = [conv, submodule, nn.ReLU(True), up, tail]
        elif innermost:
            up = nn.ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=False)
            layers = [nn.LeakyReLU(0.2, True), conv, nn.ReLU(True), up, norm_layer(outer_nc)]
        else:
            up = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=False)
            layers = [nn.LeakyReLU(0.2, True), conv, norm_layer(inner_nc), submodule, nn.ReLU(True), up,
                      norm_layer(outer_nc)]
        self.model = nn.Sequential(*layers)

    def parts(self):
        """(conv, bn_down | None, submodule | None, convT, bn_up | None)"""
        m = self.model
        if self.outermost:
            return m[0], None, m[1], m[3], None
        if self.innermost:
            return m[1], None, None, m[3], m[4]
        return m[1], m[2], m[3], m[5], m[6]

    def forward(self, x):
        raise NotImplementedError("UnetSkipConnectionBlock is executed by UnetGenerator.forward as part of the fused "
                                  "libadp_b200 U-Net; call the generator (define_G(...)) instead")


class _UnetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, gen):
        y = gen._engine_forward(x, inference=False)      # (grad mode is off inside Function.forward: say it explicitly)
        ctx.gen = gen
        ctx.token = gen._fwd_token
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        ctx.gen._engine_backward(x, y, dy, ctx.token)
        return None, None, None


class UnetGenerator(nn.Module):
    """Reference :123-152.  forward: x [B,input_nc,S,S] CUDA fp32 -> [B,1,S,S] fp32."""

    def __init__(self, cfg, input_nc, output_nc, num_downs, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        if num_downs < 5:
            raise ValueError("num_downs must be >= 5")
        kw = dict(norm_layer=norm_layer, use_dropout=use_dropout)
        block = UnetSkipConnectionBlock(cfg, ngf * 8, ngf * 8, innermost=True, norm_layer=norm_layer)
        for _ in range(num_downs - 5):
            block = UnetSkipConnectionBlock(cfg, ngf * 8, ngf * 8, submodule=block, **kw)
        for mult in (4, 2, 1):
            block = UnetSkipConnectionBlock(cfg, ngf * mult, ngf * mult * 2, submodule=block, norm_layer=norm_layer)
        self.model = UnetSkipConnectionBlock(cfg, output_nc, ngf, input_nc=input_nc, submodule=block, outermost=True,
                                             norm_layer=norm_layer)
        self.cfg = cfg
        self.input_nc, self.output_nc, self.num_downs, self.ngf = input_nc, output_nc, num_downs, ngf
        self.final_sigmoid = bool(cfg.dataset.depth_norm)
        prec = getattr(getattr(cfg, "model", None), "precision", None) or os.environ.get("ADP_PRECISION", "bf16")
        self.set_precision(prec)
        # engine state (not parameters / buffers: state_dict stays identical to the reference's)
        self._flat = None
        self._ws = None
        self._ws_key = None
        self._fwd_token = 0
        self._wcache_key = None
        self._dirty = True
        self._dirty_epoch = 0         # bumped by every mark_weights_dirty() (optimiser steps, loads, broadcasts)
        self._external_epoch = 0      # ... only by writers other than the fused optimiser
        self._mirror = None           # bf16 copy of the flat parameters kept current by FusedClipAdamW
        self._fwd_mirror = False
        self._anchor = None
        self._master_sync = None      # callable: bring the fp32 master weights up to date (sharded optimiser)
        self.grad_ready_hook = None   # callable(stage_group_index) used by the data-parallel trainer
        self.stage_groups = None      # list of (stage_begin, stage_end)
        for w in self._conv_weights():
            w.data = w.data.contiguous(memory_format=torch.channels_last)

    # ------------------------------------------------------------------ configuration
    def set_precision(self, precision):
        """'bf16': bf16 activations, tcgen05 implicit-GEMM convolutions, fp32 accumulate and fp32
        master weights.  'fp32': fp32 storage and SIMT fp32 convolutions."""
        precision = str(precision).lower()
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        self._dtype = _lib.ADP_BF16 if precision == "bf16" else _lib.ADP_F32
        return self

    def mark_weights_dirty(self, by_optimizer=False):
        self._dirty = True
        self._dirty_epoch += 1
        if not by_optimizer:          # load_state_dict, broadcasts, user code: a recorded CUDA graph must be re-recorded
            self._external_epoch += 1

    # bf16 weight mirror: the fused optimiser writes bf16(p) next to every fp32 update, so the forward pass does not
    # have to cast the 54 M weights again.  Valid only while nothing else has touched the parameters since.
    def _weights_token(self):
        return (self._dirty_epoch, self._flat["p"].data_ptr(), tuple(p._version for p in self._flat["params"]))

    def bf16_mirror(self):
        flat_p = self.flat_buffers()[0]
        if self._mirror is None or self._mirror["buf"].numel() != flat_p.numel() or self._mirror["buf"].device != flat_p.device:
            self._mirror = dict(buf=torch.empty(flat_p.numel(), device=flat_p.device, dtype=torch.bfloat16), token=None)
        return self._mirror["buf"]

    def mirror_written(self):
        """Called by the optimiser right after a step that rewrote the whole mirror."""
        self._mirror["token"] = self._weights_token()

    def _mirror_ok(self):
        return (self._dtype == _lib.ADP_BF16 and self._mirror is not None and self._mirror["token"] is not None
                and self._mirror["token"] == self._weights_token())

    # ------------------------------------------------------------------ module-tree views
    def levels(self):
        out, block = [], self.model
        while block is not None:
            conv, bn_down, sub, convT, bn_up = block.parts()
            out.append(dict(conv=conv, bn_down=bn_down, convT=convT, bn_up=bn_up))
            block = sub
        return out

    def _conv_weights(self):
        return [m.weight for lv in self.levels() for m in (lv["conv"], lv["convT"])]

    def staged_parameters(self):
        """Parameters grouped by the backward stage that finalises their gradient
        (adp_unet_backward_stages): list over stages of lists of Parameters."""
        lv, D = self.levels(), self.num_downs
        stages = []
        for s in range(2 * D):
            ps = []
            if s < D:
                ps.append(lv[s]["convT"].weight)
                if lv[s]["convT"].bias is not None:
                    ps.append(lv[s]["convT"].bias)
                if s + 1 < D and lv[s + 1]["bn_up"] is not None:
                    ps += [lv[s + 1]["bn_up"].weight, lv[s + 1]["bn_up"].bias]
            else:
                l = 2 * D - 1 - s
                ps.append(lv[l]["conv"].weight)
                if lv[l]["bn_down"] is not None:
                    ps += [lv[l]["bn_down"].weight, lv[l]["bn_down"].bias]
            stages.append(ps)
        return stages

    # ------------------------------------------------------------------ flat parameter storage
    @staticmethod
    def _view_like(flat, off, p):
        n = p.numel()
        seg = flat[off:off + n]
        if p.dim() == 4:
            a, b, kh, kw = p.shape
            return seg.view(a, kh, kw, b).permute(0, 3, 1, 2)
        return seg.view(p.shape)

    def _flat_valid(self, device):
        f = self._flat
        if f is None or f["p"].device != device:
            return False
        return all(p.data_ptr() == ptr for p, ptr in zip(f["params"], f["ptrs"]))

    @staticmethod
    def _is_bulk(p):
        """The hidden layers' convolution weights (99.97 % of the parameters): what a sharded optimiser splits."""
        return p.dim() == 4 and p.numel() >= 65536

    def _flatten(self, device):
        stages = self.staged_parameters()
        params = [p for st in stages for p in st]
        if len(params) != len(list(self.parameters())):
            raise RuntimeError("internal error: staged parameter list does not cover the module")
        # Layout: [bulk convolution weights in backward-stage order | tail: every small tensor (BatchNorm affine
        # parameters, the bias, the two thin layers' weights)].  The stage slices tile the bulk region, so a data-parallel
        # trainer reduces (or reduce-scatters) one contiguous slice per stage group and all-reduces the few-KB tail once.
        off_of, off, stage_slices = {}, 0, []
        pad = lambda n: (n + _ALIGN - 1) // _ALIGN * _ALIGN
        for st in stages:
            begin = off
            for p in st:
                if self._is_bulk(p):
                    off_of[id(p)] = off
                    off += pad(p.numel())
            stage_slices.append((begin, off))
        tail_begin = off
        for p in params:
            if not self._is_bulk(p):
                off_of[id(p)] = off
                off += pad(p.numel())
        offs = [off_of[id(p)] for p in params]
        flat_p = torch.zeros(off, device=device, dtype=torch.float32)
        flat_g = torch.zeros(off, device=device, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self._view_like(flat_p, o, p)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = None
        self._flat = dict(p=flat_p, g=flat_g, params=params, offs=offs, ptrs=[p.data_ptr() for p in params],
                          gviews=[self._view_like(flat_g, o, p) for p, o in zip(params, offs)],
                          stage_slices=stage_slices, tail=(tail_begin, off), m=None, v=None, step=0)
        for lv in self.levels():
            for bn in (lv["bn_down"], lv["bn_up"]):
                if bn is not None:
                    for name in ("running_mean", "running_var"):
                        b = getattr(bn, name)
                        if b.device != device or b.dtype != torch.float32 or not b.is_contiguous():
                            setattr(bn, name, b.to(device=device, dtype=torch.float32).contiguous())
                    if bn.num_batches_tracked.device != device:
                        bn.num_batches_tracked = bn.num_batches_tracked.to(device)
        self._anchor = torch.zeros(1, device=device, requires_grad=True)
        self._dirty = True
        self._mirror = None

    def tail_slice(self):
        """(begin, end) element offsets of the small-tensor region behind the stage slices."""
        self.flat_buffers()
        return self._flat["tail"]

    def state_dict(self, *args, **kwargs):
        if self._master_sync is not None:       # a sharded optimiser leaves other ranks' fp32 master shards stale
            self._master_sync()
        return super().state_dict(*args, **kwargs)

    def flat_buffers(self):
        """(flat parameters, flat gradients, per-stage (begin, end) element offsets of the bulk region; see tail_slice)."""
        if self._flat is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("parameters are flattened on the first CUDA forward (or once the module is on a CUDA device)")
            self._flatten(dev)
        return self._flat["p"], self._flat["g"], self._flat["stage_slices"]

    # ------------------------------------------------------------------ C structs
    def _level_array(self, use_grads, mirror=False):
        f = self._flat
        grad_of = {id(p): g for p, g in zip(f["params"], f["gviews"])} if use_grads else None
        off_of = {id(p): o for p, o in zip(f["params"], f["offs"])} if mirror else None
        base16 = self._mirror["buf"].data_ptr() if mirror else 0

        def ptr(t):
            if t is None:
                return None
            if use_grads:
                return grad_of[id(t)].data_ptr()
            return t.data_ptr()

        arr = (_lib.UnetLevel * self.num_downs)()
        for i, lv in enumerate(self.levels()):
            a = arr[i]
            a.conv_w = ptr(lv["conv"].weight)
            a.convT_w = ptr(lv["convT"].weight)
            a.convT_bias = ptr(lv["convT"].bias)
            if mirror and i > 0:
                a.conv_w_bf16 = base16 + 2 * off_of[id(lv["conv"].weight)]
                a.convT_w_bf16 = base16 + 2 * off_of[id(lv["convT"].weight)]
            for tag in ("bn_down", "bn_up"):
                bn = lv[tag]
                if bn is not None:
                    setattr(a, tag + "_w", ptr(bn.weight))
                    setattr(a, tag + "_b", ptr(bn.bias))
                    if not use_grads:
                        setattr(a, tag + "_rm", bn.running_mean.data_ptr())
                        setattr(a, tag + "_rv", bn.running_var.data_ptr())
        return arr

    def _desc(self, batch, size, training, reuse, inference=False):
        bns = [lv["bn_up"] for lv in self.levels() if lv["bn_up"] is not None]
        d = _lib.UnetDesc()
        d.batch, d.in_ch, d.out_ch, d.ngf = batch, self.input_nc, self.output_nc, self.ngf
        d.num_downs, d.size, d.dtype = self.num_downs, size, self._dtype
        d.final_sigmoid, d.training = int(self.final_sigmoid), int(training)
        d.bn_eps, d.bn_momentum = bns[0].eps, (bns[0].momentum if bns[0].momentum is not None else 0.1)
        d.reuse_weight_cache = int(reuse)
        d.inference_only = int(bool(inference) and not training)
        return d

    # ------------------------------------------------------------------ execution
    def forward(self, x):
        _lib.require_cuda(x, "UnetGenerator input", torch.float32)
        if x.dim() != 4 or x.shape[1] != self.input_nc or x.shape[2] != x.shape[3]:
            raise ValueError("expected input [B,%d,S,S], got %s" % (self.input_nc, tuple(x.shape)))
        if not self._flat_valid(x.device):
            self._flatten(x.device)
        x = x.contiguous()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._flat["params"]):
            return _UnetFunction.apply(x, self._anchor, self)
        return self._engine_forward(x, inference=True)

    def _engine_forward(self, x, inference):
        lib = _lib.load()
        B, _, S, _ = x.shape
        versions = tuple(p._version for p in self._flat["params"])
        key = (B, S, self._dtype, x.device)
        reuse = (not self.training) and (not self._dirty) and self._wcache_key == (key, versions)
        # eval mode with nothing to back-propagate: BatchNorm + activations fold into the conv epilogues
        desc = self._desc(B, S, self.training, reuse, inference=inference)
        if self._ws is None or self._ws_key != key:
            need = lib.adp_unet_workspace_bytes(ctypes.byref(desc))
            if need == 0:
                raise _lib.AdpError(lib.adp_last_error().decode())
            self._ws = torch.empty(need, device=x.device, dtype=torch.uint8)
            if os.environ.get("ADP_DEBUG_POISON", "0") == "1":      # every byte 0xFF: bf16 / fp32 / fp64 NaN patterns, so a
                self._ws.fill_(255)                                 # read of never-written workspace shows up as NaN
            self._ws_key = key
            reuse = False
            desc.reuse_weight_cache = 0
        y = torch.empty((B, self.output_nc, S, S), device=x.device, dtype=torch.float32)
        self._fwd_mirror = self._mirror_ok()
        params = self._level_array(False, self._fwd_mirror)
        with torch.cuda.device(x.device):
            _lib.check(lib.adp_unet_forward(ctypes.byref(desc), x.data_ptr(), params, self._ws.data_ptr(),
                                            self._ws.numel(), y.data_ptr(), _lib.stream_ptr()))
            if self.training:
                counters = [bn.num_batches_tracked for lv in self.levels() for bn in (lv["bn_down"], lv["bn_up"])
                            if bn is not None and bn.num_batches_tracked is not None]
                torch._foreach_add_(counters, 1)
        self._wcache_key = (key, versions) if not self.training else None
        self._dirty = False if not self.training else True
        self._fwd_token += 1
        self._last_desc = desc
        return y

    def _engine_backward(self, x, y, dy, token):
        if token != self._fwd_token:
            raise RuntimeError("UnetGenerator: the workspace of this forward pass was overwritten by a later forward; "
                               "run backward before the next forward")
        lib = _lib.load()
        f = self._flat
        desc = self._last_desc
        dy = dy.contiguous()
        if dy.dtype != torch.float32:
            dy = dy.float()
        params, grads = self._level_array(False, self._fwd_mirror), self._level_array(True)
        # stage groups exist for whoever listens to grad_ready_hook (the gradient reducer overlaps its collective with the
        # remaining stages); without a listener one call runs them all -- split-K sums are then handed from every layer to
        # its consumer (the library only does that inside one call)
        groups = (self.stage_groups if self.grad_ready_hook is not None else None) or [(0, 2 * self.num_downs)]
        with torch.cuda.device(x.device):
            for gi, (b, e) in enumerate(groups):
                _lib.check(lib.adp_unet_backward_stages(ctypes.byref(desc), x.data_ptr(), y.data_ptr(), dy.data_ptr(),
                                                        params, grads, self._ws.data_ptr(), self._ws.numel(), b, e,
                                                        _lib.stream_ptr()))
                if self.grad_ready_hook is not None:
                    self.grad_ready_hook(gi)
        for p, g in zip(f["params"], f["gviews"]):
            if p.requires_grad:
                p.grad = g

    # ------------------------------------------------------------------ state_dict compatibility
    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Accepts reference checkpoints, including DataParallel-wrapped ones whose keys carry a
        'module.' prefix (reference :52-55)."""
        if any(k.startswith("module.") for k in state_dict):
            state_dict = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.mark_weights_dirty()
        return out
