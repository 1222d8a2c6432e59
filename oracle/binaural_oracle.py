"""Plain-PyTorch fp32 restatement of the reference's BinauralAttentionDepthNet forward (BASELINE config 4), as a function
of a state_dict.

Test infrastructure only (see oracle/__init__.py).  Parity: pinned against the unmodified
models/binaural_attention_model.py by tests/test_oracle_golden.py (tests/golden/binaural.npz, oracle/gen_golden.py
gen_binaural): forward, autograd gradients and the eval forward.

Follows models/binaural_attention_model.py:
* DoubleConv :22-39      conv3x3(p=1, no bias) -> BN -> ReLU, twice
* Down :42-53, Up :56-78 MaxPool2d(2); bilinear x2 (align_corners=True) or, bilinear=False, ConvTranspose2d(k2, s2) (taken
  when the state_dict holds up{i}.up.weight); pad, cat([skip, up]), DoubleConv(in, out, in // 2 | out)
* BinauralCrossAttention :81-153   shared q/k/v/out 1x1 convs, softmax(q^T k / sqrt(C)), residual scaled by gamma
* BinauralAttentionDepthNet.forward :279-334   two encoders, attention at `attention_levels`, fusion 1x1 + BN + ReLU,
  decoder, sigmoid head * max_depth, (interpolate to output_size), clamp
"""
import torch
import torch.nn.functional as F


def init_state_dict(base_channels=64, attention_levels=(2, 3, 4, 5), seed=0):
    """Deterministic weights with the reference's shapes and key order (kaiming fan_out convs, BN gamma 1 / beta 0 as
    :268-277, but from numpy-free torch generators so that it does not depend on module construction order)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, bias):
        fan_out = cout * k * k
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (2.0 / fan_out) ** 0.5
        if bias:
            sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    def double(name, cin, cout, mid=None):
        mid = mid or cout
        conv(name + ".double_conv.0", mid, cin, 3, False); bn(name + ".double_conv.1", mid)
        conv(name + ".double_conv.3", cout, mid, 3, False); bn(name + ".double_conv.4", cout)

    b = base_channels
    ch = {1: b, 2: 2 * b, 3: 4 * b, 4: 8 * b, 5: 8 * b}
    for enc in ("left_encoder", "right_encoder"):
        double(enc + ".inc", 1, b)
        for i, (cin, cout) in enumerate(((b, 2 * b), (2 * b, 4 * b), (4 * b, 8 * b), (8 * b, 8 * b)), 1):
            double("%s.down%d.maxpool_conv.1" % (enc, i), cin, cout)
    for lv in attention_levels:
        c = ch[lv]
        stem = "attention_modules.attn_%d" % lv
        sd[stem + ".gamma"] = torch.full((1,), 0.5)
        conv(stem + ".query", c // 8, c, 1, True); conv(stem + ".key", c // 8, c, 1, True)
        conv(stem + ".value", c, c, 1, True); conv(stem + ".out", c, c, 1, True)
    for lv in (1, 2, 3, 4, 5):
        conv("fusion_layers.fusion_%d.0" % lv, ch[lv], 2 * ch[lv], 1, True); bn("fusion_layers.fusion_%d.1" % lv, ch[lv])
    double("up1.conv", 16 * b, 4 * b, 8 * b); double("up2.conv", 8 * b, 2 * b, 4 * b)
    double("up3.conv", 4 * b, b, 2 * b); double("up4.conv", 2 * b, b, b)
    conv("outc.0", 1, b, 1, True)
    sd["outc.0.weight"] = sd["outc.0.weight"] * 0.1        # keep the sigmoid head of the untrained network unsaturated
    sd["outc.0.bias"] = torch.full((1,), -0.3)
    return sd


def _bn_relu(x, sd, stem, training, update):
    rm, rv = sd[stem + ".running_mean"], sd[stem + ".running_var"]
    if training and not update:
        rm, rv = rm.clone(), rv.clone()
    return F.relu(F.batch_norm(x, rm, rv, sd[stem + ".weight"], sd[stem + ".bias"], training, 0.1, 1e-5))


def _double(x, sd, stem, training, update):
    x = _bn_relu(F.conv2d(x, sd[stem + ".double_conv.0.weight"], padding=1), sd, stem + ".double_conv.1", training, update)
    return _bn_relu(F.conv2d(x, sd[stem + ".double_conv.3.weight"], padding=1), sd, stem + ".double_conv.4", training, update)


def _attend(a, b, sd, stem):
    B, C, H, W = a.shape
    q = F.conv2d(a, sd[stem + ".query.weight"], sd[stem + ".query.bias"]).view(B, -1, H * W)
    k = F.conv2d(b, sd[stem + ".key.weight"], sd[stem + ".key.bias"]).view(B, -1, H * W)
    v = F.conv2d(b, sd[stem + ".value.weight"], sd[stem + ".value.bias"]).view(B, C, H * W)
    att = torch.softmax(torch.bmm(q.transpose(1, 2), k) / (C ** 0.5), dim=-1)
    out = torch.bmm(v, att.transpose(1, 2)).view(B, C, H, W)
    out = F.conv2d(out, sd[stem + ".out.weight"], sd[stem + ".out.bias"])
    return a + sd[stem + ".gamma"] * out


def forward(sd, x, attention_levels=(2, 3, 4, 5), max_depth=30.0, output_size=None, training=True, update_running=False):
    """x [B,2,H,W] fp32 -> depth [B,1,H,W].  sd: reference state_dict (tensors may require grad)."""
    feats = []
    for enc, ch in (("left_encoder", 0), ("right_encoder", 1)):
        f = {1: _double(x[:, ch:ch + 1], sd, enc + ".inc", training, update_running)}
        for i in (1, 2, 3, 4):
            f[i + 1] = _double(F.max_pool2d(f[i], 2), sd, "%s.down%d.maxpool_conv.1" % (enc, i), training, update_running)
        feats.append(f)
    left, right = feats
    fused = {}
    for lv in (1, 2, 3, 4, 5):
        lf, rf = left[lv], right[lv]
        if lv in attention_levels:
            stem = "attention_modules.attn_%d" % lv
            lf, rf = _attend(lf, rf, sd, stem), _attend(rf, lf, sd, stem)
        stem = "fusion_layers.fusion_%d" % lv
        e = F.conv2d(torch.cat([lf, rf], 1), sd[stem + ".0.weight"], sd[stem + ".0.bias"])
        fused[lv] = _bn_relu(e, sd, stem + ".1", training, update_running)
    y = fused[5]
    for name, skip in (("up1", 4), ("up2", 3), ("up3", 2), ("up4", 1)):
        if name + ".up.weight" in sd:        # bilinear=False (:65-66): ConvTranspose2d(in, in // 2, kernel_size=2, stride=2)
            y = F.conv_transpose2d(y, sd[name + ".up.weight"], sd[name + ".up.bias"], stride=2)
        else:
            y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
        s = fused[skip]
        dy_, dx_ = s.shape[2] - y.shape[2], s.shape[3] - y.shape[3]
        y = F.pad(y, [dx_ // 2, dx_ - dx_ // 2, dy_ // 2, dy_ - dy_ // 2])
        y = _double(torch.cat([s, y], 1), sd, name + ".conv", training, update_running)
    depth = torch.sigmoid(F.conv2d(y, sd["outc.0.weight"], sd["outc.0.bias"])) * max_depth
    if output_size is not None and depth.shape[-1] != output_size:
        depth = F.interpolate(depth, size=(output_size, output_size), mode="bilinear", align_corners=False)
    return torch.clamp(depth, 0, max_depth)
