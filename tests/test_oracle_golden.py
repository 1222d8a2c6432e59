"""Pin the CPU oracle (oracle/) against vectors produced by the unmodified
reference modules (oracle/gen_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from audio_depth_estimation_b200 import synthetic
from oracle import feature_oracle as fo
from oracle import loss_oracle as lo
from oracle import unet_oracle as uo


def rel_to_max(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def feat(golden_dir):
    return np.load(os.path.join(golden_dir, "feature.npz"))


def test_stft_small_512(feat):
    w = synthetic.waveform(1, 1000, seed=13)[0]
    got = fo.stft_mag(w, 512, 64, 16)
    assert got.shape == feat["small_spec_512"].shape
    assert rel_to_max(got, feat["small_spec_512"]) <= 1e-5


def test_stft_small_400(feat):
    w = synthetic.waveform(1, 2000, seed=14)[0]
    got = fo.stft_mag(w, 400, 200, 100)
    assert got.shape == feat["small_spec_400"].shape
    assert rel_to_max(got, feat["small_spec_400"]) <= 1e-5


@pytest.mark.parametrize("name,echo", [("v2", False), ("v2echo", True)])
def test_feature_v2(feat, name, echo):
    w = synthetic.waveform(1, 8000, seed=11, echo=echo)[0]
    spec = fo.stft_mag(w[:, :fo.cut_length(30.0)], 512, 64, 16)
    assert spec.shape == (2, 257, 487)
    assert rel_to_max(spec[:, :, ::37], feat[name + "_spec_slice"]) <= 1e-5
    got = fo.feature_v2(w, 30.0, 256)
    # After log + min-max the values are in [0,1].  The reference's own fp32 FFT
    # is only accurate to ~7e-4 *element-relative* on near-zero bins (SURVEY.md
    # section 7), the log turns that into an absolute error and the per-channel
    # min (taken at exactly such a bin) shifts the whole plane, so the feature
    # is compared with an absolute bound of 5e-4 (2e-3 for the echo variant,
    # whose dynamic range is far wider); the spectrogram itself is held to 1e-5
    # of its max above.
    assert np.abs(got - feat[name + "_feat"]).max() <= (2e-3 if echo else 5e-4)


def test_feature_v1(feat):
    w = synthetic.waveform(1, synthetic.V1_LEN, seed=12)[0]
    spec = fo.stft_mag(w, 512, 64, 16)
    assert spec.shape == (2, 257, 201)
    assert rel_to_max(spec[:, :, ::13], feat["v1_spec_slice"]) <= 1e-5
    assert rel_to_max(fo.feature_v1(w), feat["v1_feat"]) <= 1e-5


def test_resize(feat):
    rng = np.random.default_rng(15)
    plane = rng.uniform(0, 1, size=(2, 257, 101)).astype(np.float32)
    assert np.abs(fo.resize(plane, 64) - feat["resize_257x101_to_64"]).max() <= 1e-6


def test_loss(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    for i, (shape, dn, md) in enumerate([((2, 1, 64, 64), False, 30.0), ((3, 1, 32, 32), True, 12.0)]):
        gt = synthetic.gt_depth(shape[0], shape[2], md, seed=710 + i, normalised=dn)
        pred = g["case%d_pred" % i]
        loss, l1, si, grad = lo.depth_loss_and_grad(pred, gt, scale=md if dn else 1.0)
        ref = g["case%d_loss" % i]
        assert abs(loss - ref[0]) <= 1e-5 * abs(ref[0])
        assert abs(l1 - ref[1]) <= 1e-5 * abs(ref[1])
        assert abs(si - ref[2]) <= 1e-5 * abs(ref[2])
        rg = g["case%d_grad" % i]
        assert np.abs(grad - rg).max() <= 1e-5 * np.abs(rg).max()
        # the torch restatement used for the U-Net step agrees too
        tl = uo.depth_loss(torch.from_numpy(pred), torch.from_numpy(gt), depth_norm=dn, max_depth=md)
        assert abs(tl.item() - ref[0]) <= 1e-5 * abs(ref[0])


UNET_CASES = {
    "u128_ngf16_b3_sigmoid": ("unet_128", 16, 3, 128, True, 12.0, 200, True, False),
    "u128_ngf64_b2_relu": ("unet_128", 64, 2, 128, False, 30.0, 300, True, False),
    "u128_ngf16_b2_eval": ("unet_128", 16, 2, 128, False, 30.0, 400, False, True),
    "u256_ngf64_b2_relu": ("unet_256", 64, 2, 256, False, 30.0, 100, True, False),
    "u256_ngf64_b1_eval": ("unet_256", 64, 1, 256, True, 12.0, 500, False, True),
}


def build_case(case):
    netG, ngf, batch, size, dn, md, seed, train, warm = case
    nd = 8 if netG == "unet_256" else 7
    sd = uo.make_state_dict(ngf, nd, seed=seed)
    if warm:
        rng = np.random.default_rng(seed + 1)
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = torch.from_numpy(rng.normal(0, 0.05, sd[k].shape).astype(np.float32))
            if k.endswith("running_var"):
                sd[k] = torch.from_numpy(rng.uniform(0.5, 1.5, sd[k].shape).astype(np.float32))
    x = torch.from_numpy(synthetic.feature_like(batch, size, seed=seed + 2))
    gt = torch.from_numpy(synthetic.gt_depth(batch, size, md, seed=seed + 3, normalised=dn))
    return nd, sd, x, gt


@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_oracle(golden_dir, name):
    case = UNET_CASES[name]
    netG, ngf, batch, size, dn, md, seed, train, warm = case
    g = np.load(os.path.join(golden_dir, "unet_%s.npz" % name))
    nd, sd, x, gt = build_case(case)
    torch.set_num_threads(8)
    if not train:
        with torch.no_grad():
            y = uo.unet_forward(x, sd, nd, dn, training=False)
        assert rel_to_max(y.numpy(), g["y"]) <= 1e-5
        return
    names = [str(n) for n in g["param_names"]]
    for n in names:
        sd["model." + n if not n.startswith("model.") else n].requires_grad_(True)
    y = uo.unet_forward(x, sd, nd, dn, training=True)
    y.retain_grad()
    loss = uo.depth_loss(y, gt, depth_norm=dn, max_depth=md)
    loss.backward()
    assert rel_to_max(y.detach().numpy(), g["y"]) <= 1e-4
    assert abs(loss.item() - g["loss"][0]) <= 1e-4 * abs(g["loss"][0])
    assert rel_to_max(y.grad.numpy(), g["dy"]) <= 1e-4
    gn = np.array([sd[n].grad.double().norm().item() for n in names])
    assert np.all(np.abs(gn - g["grad_norms"]) <= 2e-3 * g["grad_norms"].max())
    stats = [str(s) for s in g["stat_names"]]
    got = np.stack([sd[s][:8].numpy() for s in stats])
    assert np.abs(got - g["stat_head"]).max() <= 1e-4
    # optimiser restatement
    params = [sd[n].detach() for n in names]
    grads = [sd[n].grad for n in names]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    tn = uo.clip_adamw_step(params, grads, m, v, 1, 0.002)
    assert abs(tn.item() - g["total_norm"][0]) <= 1e-3 * g["total_norm"][0]
    head = np.stack([p.reshape(-1)[:16].numpy() if p.numel() >= 16 else
                     np.pad(p.reshape(-1).numpy(), (0, 16 - p.numel())) for p in params])
    assert np.abs(head - g["param_head_after"]).max() <= 2e-4


# ---- mel branch ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def melgold(golden_dir):
    return np.load(os.path.join(golden_dir, "feature_mel.npz"))


def test_mel_fbanks_match_torchaudio():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(257, 20.0, 20000.0, 32, 44100, norm=None, mel_scale="htk").numpy()
    got = fo.mel_fbanks(257, 20.0, 20000.0, 32, 44100)
    # float32 cancellation in f_pts - all_freqs: weights agree to a few 1e-6 absolute (peak weight is 1)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-5


@pytest.mark.parametrize("name,echo,seed", [("mel", False, 21), ("melecho", True, 22)])
def test_feature_v2_mel(melgold, name, echo, seed):
    w = synthetic.waveform(1, 8000, seed=seed, echo=echo)[0]
    spec = fo.mel_spectrogram(w[:, :fo.cut_length(30.0)], 512, 64)
    assert spec.shape == (2, 32, 244)
    assert rel_to_max(spec, melgold[name + "_spec"]) <= 1e-5
    got = fo.feature_v2_mel(w, 30.0, 256)
    # a mel band sums >= 1 magnitude bins, so the near-zero-bin problem of the linear branch does not arise
    assert np.abs(got - melgold[name + "_feat"]).max() <= 5e-5


def test_mel_defaults_400(melgold):
    w = synthetic.waveform(1, 3000, seed=23)[0]
    got = fo.mel_spectrogram(w)                      # n_fft 400, win 400, hop 200
    assert got.shape == melgold["small_mel_400"].shape
    assert rel_to_max(got, melgold["small_mel_400"]) <= 1e-5


# ---- evaluation metrics -------------------------------------------------------------------------
METRIC_CASES = [("m30", False, 30.0, 64, 0), ("n12", True, 12.0, 48, 1), ("zeros", False, 30.0, 32, 2),
                ("negpred", False, 30.0, 32, 3), ("emptygt", False, 30.0, 32, 4)]


def metric_case_inputs(g, name, dn, md, size, k):
    gt = synthetic.gt_depth(3, size, md, seed=910 + k, normalised=dn)
    if name == "emptygt":
        gt[1] = 0.0
    return gt, g[name + "_pred"]


@pytest.mark.parametrize("case", METRIC_CASES, ids=[c[0] for c in METRIC_CASES])
def test_metrics_oracle(golden_dir, case):
    from oracle import metrics_oracle as mo
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    name, dn, md, size, k = case
    gt, pred = metric_case_inputs(g, *case)
    got = mo.batch_errors(gt, pred, dn, md)
    ref = g[name + "_errors"]
    # the reference evaluates in float32 numpy; the oracle in float64
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_metrics_oracle_raw_branches(golden_dir):
    from oracle import metrics_oracle as mo
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    gt, pred = g["raw_gt"], g["raw_pred"]
    for key, (a, b) in {"raw": (gt, pred), "rawneg": (gt, -np.abs(pred) - 1.0),
                        "rawsmall": (gt / 100.0, np.abs(pred) / 100.0)}.items():
        got = np.array(mo.compute_errors(a, b))
        ref = g[key + "_errors"]
        assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), key


# ---- config 4: BinauralAttentionDepthNet -----------------------------------------------------------------------
@pytest.mark.parametrize("name,levels", [("lv345_b2", (3, 4, 5)), ("lv2345_b2", (2, 3, 4, 5))])
def test_binaural_oracle_matches_reference(golden_dir, name, levels):
    """oracle/binaural_oracle.forward on the reference's own initial weights (same seed, same module construction order
    in the mirror) reproduces the reference's forward, gradients and eval forward."""
    from audio_depth_estimation_b200.models.binaural_attention_model import BinauralAttentionDepthNet
    from oracle import binaural_oracle as bo
    g = np.load(os.path.join(golden_dir, "binaural.npz"))
    torch.manual_seed(0)
    net = BinauralAttentionDepthNet(64, True, 128, 30.0, list(levels))      # parameter container only (CPU)
    with torch.no_grad():
        for m in net.attention_modules.values():
            m.gamma.fill_(0.5)
        net.outc[0].weight.mul_(0.1)
        net.outc[0].bias.fill_(-1.2)
    sd = {k: v.detach().clone().contiguous() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    x = torch.from_numpy(synthetic.feature_like(2, 128, seed=301))
    r = torch.from_numpy(np.random.default_rng(302).normal(0, 1, (2, 1, 128, 128)).astype(np.float32))
    torch.set_num_threads(8)
    y = bo.forward(sd, x, levels, 30.0, training=True, update_running=True)
    assert rel_to_max(y.detach().numpy(), g[name + "_y"]) <= 1e-4
    (y * r).sum().backward()
    names = list(g[name + "_grad_names"])
    for i, k in enumerate(names):
        ref = g[name + "_grad_norms"][i]
        if ref < 1e-4 * g[name + "_grad_norms"].max():
            continue
        assert abs(float(sd[k].grad.double().norm()) - ref) <= 2e-3 * ref, k
    assert np.abs(sd["fusion_layers.fusion_3.1.running_mean"].numpy() - g[name + "_rm_fusion3"]).max() <= 1e-4
    with torch.no_grad():
        ye = bo.forward(sd, x, levels, 30.0, training=False)
    assert rel_to_max(ye.numpy(), g[name + "_y_eval"]) <= 1e-4


def test_binaural_oracle_transposed_conv_decoder_and_resize(golden_dir):
    """bilinear=False + output_size != input size (reference :65-66, :322-328): oracle/binaural_oracle.forward on the
    reference's own initial weights against tests/golden/binaural_ct.npz (gen_golden.py gen_binaural_ct)."""
    from audio_depth_estimation_b200.models.binaural_attention_model import BinauralAttentionDepthNet
    from oracle import binaural_oracle as bo
    g = np.load(os.path.join(golden_dir, "binaural_ct.npz"))
    torch.manual_seed(0)
    net = BinauralAttentionDepthNet(64, False, 96, 30.0, [4, 5])      # parameter container only (CPU)
    with torch.no_grad():
        for m in net.attention_modules.values():
            m.gamma.fill_(0.5)
        net.outc[0].weight.mul_(0.1)
        net.outc[0].bias.fill_(-1.2)
    sd = {k: v.detach().clone().contiguous() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    x = torch.from_numpy(synthetic.feature_like(2, 128, seed=311))
    r = torch.from_numpy(np.random.default_rng(312).normal(0, 1, (2, 1, 96, 96)).astype(np.float32))
    torch.set_num_threads(8)
    y = bo.forward(sd, x, (4, 5), 30.0, output_size=96, training=True, update_running=True)
    assert y.shape == (2, 1, 96, 96) and rel_to_max(y.detach().numpy(), g["y"]) <= 1e-4
    (y * r).sum().backward()
    norms = g["grad_norms"]
    for i, k in enumerate(list(g["grad_names"])):
        if norms[i] < 1e-4 * norms.max():
            continue
        assert abs(float(sd[k].grad.double().norm()) - norms[i]) <= 2e-3 * norms[i], k
    with torch.no_grad():
        ye = bo.forward(sd, x, (4, 5), 30.0, output_size=96, training=False)
    assert rel_to_max(ye.numpy(), g["y_eval"]) <= 1e-4
