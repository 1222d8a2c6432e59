"""Plain-PyTorch fp32 restatement of the reference U-Net dataflow, the training
step around it, and the state_dict key map.

Test infrastructure only (see oracle/__init__.py).  Parity: pinned against the
unmodified reference modules by tests/test_oracle_golden.py (vectors made by
oracle/gen_golden.py).

Follows models/unetbaseline_model.py:
* UnetGenerator.__init__ :123-148   (nesting, channel plan)
* UnetSkipConnectionBlock :157-235  (down = [LeakyReLU(0.2,inplace)] Conv(k4,s2,p1)
  [BN]; up = ReLU(inplace) ConvT(k4,s2,p1) [BN]; outermost tail ReLU | Sigmoid
  chosen by cfg.dataset.depth_norm :201-206; cat([x, model(x)]) :231-235)
* the in-place LeakyReLU makes the concatenated skip LeakyReLU(e), and the
  parent's in-place ReLU then turns it into ReLU(e).
The step body follows train.py:633-693.
"""
import numpy as np
import torch
import torch.nn.functional as F


def channel_plan(ngf, num_downs):
    """Encoder output channels per level, level 0 = outermost conv.
    unetbaseline_model.py:140-148."""
    ch = [ngf, ngf * 2, ngf * 4, ngf * 8]
    ch += [ngf * 8] * (num_downs - 4)
    return ch[:num_downs] if num_downs >= 4 else None


def level_keys(num_downs, prefix="model."):
    """state_dict key stems per level.  Returns a list of
    dicts with keys conv, bn_down (or None), convT, convT_bias (or None),
    bn_up (or None)."""
    out = []
    stem = prefix + "model"          # outermost Sequential
    for lvl in range(num_downs):
        if lvl == 0:
            out.append(dict(conv=stem + ".0.weight", bn_down=None,
                            convT=stem + ".3.weight", convT_bias=stem + ".3.bias",
                            bn_up=None))
            stem = stem + ".1.model"
        elif lvl < num_downs - 1:
            out.append(dict(conv=stem + ".1.weight", bn_down=stem + ".2",
                            convT=stem + ".5.weight", convT_bias=None,
                            bn_up=stem + ".6"))
            stem = stem + ".3.model"
        else:
            out.append(dict(conv=stem + ".1.weight", bn_down=None,
                            convT=stem + ".3.weight", convT_bias=None,
                            bn_up=stem + ".4"))
    return out


def _bn(x, sd, stem, training, eps, momentum, update_running):
    rm, rv = sd[stem + ".running_mean"], sd[stem + ".running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[stem + ".weight"], sd[stem + ".bias"],
                     training, momentum, eps)
    if training and update_running and (stem + ".num_batches_tracked") in sd:
        sd[stem + ".num_batches_tracked"] += 1
    return y


def unet_forward(x, sd, num_downs=8, depth_norm=False, training=True,
                 eps=1e-5, momentum=0.1, update_running=True, prefix="model.",
                 return_intermediates=False):
    """x [B,in,H,W] fp32 -> [B,out,H,W].  sd: dict name -> tensor in the
    reference's state_dict layout (parameters may require grad)."""
    keys = level_keys(num_downs, prefix)
    acts = []           # a_k (activated encoder outputs, k = 1..num_downs-1)
    inter = {}
    h = x
    for lvl, k in enumerate(keys):
        if lvl > 0:
            h = F.leaky_relu(h, 0.2)
            acts.append(h)          # the skip the decoder sees (in-place quirk)
        h = F.conv2d(h, sd[k["conv"]], None, stride=2, padding=1)
        if k["bn_down"] is not None:
            h = _bn(h, sd, k["bn_down"], training, eps, momentum, update_running)
        inter["e%d" % (lvl + 1)] = h
    # decoder, innermost first
    for lvl in range(num_downs - 1, -1, -1):
        k = keys[lvl]
        if lvl < num_downs - 1:
            h = torch.cat([acts[lvl], h], 1)
        h = F.relu(h)
        bias = sd[k["convT_bias"]] if k["convT_bias"] is not None else None
        h = F.conv_transpose2d(h, sd[k["convT"]], bias, stride=2, padding=1)
        if k["bn_up"] is not None:
            h = _bn(h, sd, k["bn_up"], training, eps, momentum, update_running)
        inter["u%d" % (lvl + 1)] = h
    y = torch.sigmoid(h) if depth_norm else F.relu(h)
    if return_intermediates:
        return y, inter
    return y


def depth_loss(pred, gt, criterion="Combined", l1_weight=0.237, silog_weight=0.637,
               silog_lambda=0.869, depth_norm=False, max_depth=30.0, eps=1e-6):
    """Mask + criterion wiring of train.py:646-669 with utils_loss.py:29-49."""
    mask = gt != 0.0
    p, g = pred[mask], gt[mask]
    if depth_norm:
        p, g = p * max_depth, g * max_depth

    def silog(p, g, lam):
        pc, gc = torch.clamp(p, min=eps), torch.clamp(g, min=eps)
        d = torch.log(pc) - torch.log(gc)
        v = torch.mean(d ** 2) - lam * (torch.mean(d) ** 2)
        return torch.sqrt(torch.clamp(v, min=0.0))

    if criterion == "L1":
        return torch.mean(torch.abs(p - g))
    if criterion == "SIlog":
        return silog(p, g, silog_lambda)
    if criterion == "Combined":
        loss = l1_weight * torch.mean(torch.abs(p - g))
        if silog_weight != 0.0:
            loss = loss + silog_weight * silog(p, g, silog_lambda)
        return loss
    raise ValueError(criterion)


def make_state_dict(ngf=64, num_downs=8, in_ch=2, out_ch=1, seed=0, prefix="model.",
                    dtype=torch.float32):
    """Deterministic weights in the reference's state_dict layout, drawn with a
    numpy generator so they are identical on every machine (the distribution is
    init_weights('normal', 0.02): conv ~ N(0,0.02), BN weight ~ N(1,0.02), BN
    bias 0 -- unetbaseline_model.py:9-40; BN bias is drawn ~N(0,0.02) here so
    that parity tests exercise it)."""
    rng = np.random.default_rng(seed)
    ch = channel_plan(ngf, num_downs)
    sd = {}

    def normal(shape, mean, std):
        return torch.from_numpy((rng.standard_normal(shape) * std + mean).astype(np.float32)).to(dtype)

    def bn(stem, c):
        sd[stem + ".weight"] = normal((c,), 1.0, 0.02)
        sd[stem + ".bias"] = normal((c,), 0.0, 0.02)
        sd[stem + ".running_mean"] = torch.zeros(c, dtype=dtype)
        sd[stem + ".running_var"] = torch.ones(c, dtype=dtype)
        sd[stem + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)

    keys = level_keys(num_downs, prefix)
    for lvl, k in enumerate(keys):
        cin = in_ch if lvl == 0 else ch[lvl - 1]
        cout = ch[lvl]
        sd[k["conv"]] = normal((cout, cin, 4, 4), 0.0, 0.02)
        if k["bn_down"] is not None:
            bn(k["bn_down"], cout)
        # convT: input = inner (x2 when it sees a concat), output = outer
        t_in = cout if lvl == num_downs - 1 else cout * 2
        t_out = out_ch if lvl == 0 else ch[lvl - 1]
        sd[k["convT"]] = normal((t_in, t_out, 4, 4), 0.0, 0.02)
        if k["convT_bias"] is not None:
            sd[k["convT_bias"]] = normal((t_out,), 0.0, 0.02)
        if k["bn_up"] is not None:
            bn(k["bn_up"], t_out)
    return sd


def ordered_state_dict(sd, num_downs, prefix="model."):
    """Re-order to the reference's state_dict() iteration order."""
    order = []

    def rec(lvl):
        k = level_keys(num_downs, prefix)[lvl]
        order.append(k["conv"])
        if k["bn_down"]:
            order.extend(k["bn_down"] + s for s in
                         (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked"))
        if lvl < num_downs - 1:
            rec(lvl + 1)
        order.append(k["convT"])
        if k["convT_bias"]:
            order.append(k["convT_bias"])
        if k["bn_up"]:
            order.extend(k["bn_up"] + s for s in
                         (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked"))
    rec(0)
    return {k: sd[k] for k in order}


def clip_adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr, max_norm=1.0,
                    betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """clip_grad_norm_(max_norm) + torch.optim.AdamW defaults (train.py:471-476,
    :689-691), restated on lists of tensors, in place.  Returns the total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        p.mul_(1.0 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)
    return total
