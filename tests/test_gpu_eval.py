"""GPU parity for the rows SURVEY.md 8(f) lists next: the mel branch of the V2 feature transform and the
validation metrics (compute_errors) -- both through the C ABI, against golden vectors of the unmodified reference
and against the oracle on seeded inputs."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from audio_depth_estimation_b200 import synthetic

pytestmark = pytest.mark.gpu


def rel_to_max(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def melgold(golden_dir):
    return np.load(os.path.join(golden_dir, "feature_mel.npz"))


@pytest.fixture(scope="module")
def metgold(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics.npz"))


@pytest.mark.parametrize("tc", [1, 0], ids=["tc", "simt"])
@pytest.mark.parametrize("name,echo,seed", [("mel", False, 21), ("melecho", True, 22)])
def test_mel_feature_golden(melgold, name, echo, seed, tc):
    from audio_depth_estimation_b200 import _lib, feature
    _lib.load().adp_set_tensor_core(tc)
    try:
        w = torch.from_numpy(synthetic.waveform(1, 8000, seed=seed, echo=echo)[0]).cuda()
        spec = feature.melspectrogram(w, n_fft=512, win_length=64, length=feature.cut_length(30.0))
        assert spec.shape == (2, 32, 244)
        assert rel_to_max(spec.cpu().numpy(), melgold[name + "_spec"]) <= 1e-4         # north_star: <= 1e-4 of max
        cfg = SimpleNamespace(dataset=SimpleNamespace(name="batvisionv2", images_size=256, max_depth=30.0,
                                                      audio_format="mel_spectrogram"))
        feat = feature.SpectrogramTransform.for_cfg(cfg)(w[None])[0]
        assert np.abs(feat.cpu().numpy() - melgold[name + "_feat"]).max() <= 5e-4
    finally:
        _lib.load().adp_set_tensor_core(1)


def test_mel_defaults_golden(melgold):
    from audio_depth_estimation_b200 import feature
    w = torch.from_numpy(synthetic.waveform(1, 3000, seed=23)[0]).cuda()
    got = feature.melspectrogram(w)                       # n_fft 400, win 400, hop 200 (generic STFT kernel)
    assert got.shape == (2, 32, 16) and rel_to_max(got.cpu().numpy(), melgold["small_mel_400"]) <= 1e-4


def test_mel_batch_vs_oracle():
    from audio_depth_estimation_b200 import feature
    from oracle import feature_oracle as fo
    w = synthetic.waveform(5, 8200, seed=31)
    t = feature.SpectrogramTransform(128, 30.0, mel={"n_mels": 32})
    got = t(torch.from_numpy(w).cuda()).cpu().numpy()
    assert got.shape == (5, 2, 128, 128)
    for b in range(5):
        assert np.abs(got[b] - fo.feature_v2_mel(w[b], 30.0, 128)).max() <= 5e-4
    # other bank geometries go through the same kernel
    m = feature.melspectrogram(torch.from_numpy(w[0]).cuda(), n_fft=512, win_length=64, n_mels=48, f_min=50.0, f_max=16000.0)
    ref = fo.mel_spectrogram(w[0], 512, 64, f_min=50.0, f_max=16000.0, n_mels=48)
    assert rel_to_max(m.cpu().numpy(), ref) <= 1e-4


METRIC_CASES = [("m30", False, 30.0, 64, 0), ("n12", True, 12.0, 48, 1), ("zeros", False, 30.0, 32, 2),
                ("negpred", False, 30.0, 32, 3), ("emptygt", False, 30.0, 32, 4)]


@pytest.mark.parametrize("case", METRIC_CASES, ids=[c[0] for c in METRIC_CASES])
def test_metrics_golden(metgold, case):
    from audio_depth_estimation_b200 import utils_criterion as uc
    name, dn, md, size, k = case
    gt = synthetic.gt_depth(3, size, md, seed=910 + k, normalised=dn)
    if name == "emptygt":
        gt[1] = 0.0
    cfg = SimpleNamespace(dataset=SimpleNamespace(depth_norm=dn, max_depth=md))
    got = uc.batch_errors(torch.from_numpy(gt).cuda(), torch.from_numpy(metgold[name + "_pred"]).cuda(), cfg)
    ref = metgold[name + "_errors"]
    assert got.shape == (3, 7) and got.dtype == torch.float64
    assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    # the threshold counts are integer work: exact
    assert np.allclose(got.cpu().numpy()[:, 2:5], ref[:, 2:5], atol=1e-7)


def test_compute_errors_branches_golden(metgold):
    from audio_depth_estimation_b200 import utils_criterion as uc
    gt, pred = metgold["raw_gt"], metgold["raw_pred"]
    for key, (a, b) in {"raw": (gt, pred), "rawneg": (gt, -np.abs(pred) - 1.0),
                        "rawsmall": (gt / 100.0, np.abs(pred) / 100.0)}.items():
        got = np.array(uc.compute_errors(torch.from_numpy(a).cuda(), torch.from_numpy(b.astype(np.float32)).cuda()))
        ref = metgold[key + "_errors"]
        assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), key
    assert uc.compute_errors(torch.zeros(8, 8).cuda(), torch.ones(8, 8).cuda()) == (0.0,) * 7


def test_metrics_batch_vs_oracle_full_size():
    from audio_depth_estimation_b200 import utils_criterion as uc
    from oracle import metrics_oracle as mo
    rng = np.random.default_rng(5)
    gt = synthetic.gt_depth(16, 256, 30.0, seed=41, normalised=False)
    pred = (gt + rng.normal(0, 3.0, gt.shape)).astype(np.float32)
    pred[3] = 1e-7                                           # fall-back set: 0 < pred <= eps everywhere
    pred[5] = -2.0                                           # clipped up to eps by the preparation
    got = uc.batch_errors(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda(), depth_norm=False, max_depth=30.0)
    ref = mo.batch_errors(gt, pred, False, 30.0)
    assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_metrics_test_protocol_and_eval_loop():
    """test.py:262-276 protocol (clip at 0 only) and the evaluation loop with one D2H at the end."""
    from audio_depth_estimation_b200 import utils_criterion as uc
    from audio_depth_estimation_b200.config_loader import load_config
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    from audio_depth_estimation_b200.training import TrainStep, evaluate_loader
    from oracle import metrics_oracle as mo
    rng = np.random.default_rng(6)
    gt = synthetic.gt_depth(4, 64, 12.0, seed=43, normalised=True)
    pred = (gt + rng.normal(0, 0.2, gt.shape)).astype(np.float32)      # plenty of negative predictions
    got = uc.batch_errors(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda(), depth_norm=True, max_depth=12.0,
                          protocol="test")
    ref = mo.batch_errors(gt, pred, True, 12.0, protocol="test")
    assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())

    cfg = load_config()
    cfg.dataset.images_size, cfg.model.generator, cfg.model.precision = 128, "unet_128", "fp32"
    torch.manual_seed(0)
    net = define_G(cfg, 2, 1, 16, "unet_128", "batch", False, gpu_ids=[0])
    step = TrainStep(cfg, net, waveform_input=False)
    batches = []
    for i in range(3):
        x = torch.from_numpy(synthetic.feature_like(2, 128, seed=50 + i)).cuda()
        g = torch.from_numpy(synthetic.gt_depth(2, 128, cfg.dataset.max_depth, seed=60 + i,
                                                normalised=bool(cfg.dataset.depth_norm))).cuda()
        batches.append((x, g))
    res = evaluate_loader(step, batches, protocol="test")
    tabs, losses = [], []
    for x, g in batches:
        y, loss = step.evaluate(x, g)
        losses.append(float(loss))
        tabs.append(mo.batch_errors(g.cpu().numpy(), y.cpu().numpy(), bool(cfg.dataset.depth_norm),
                                    float(cfg.dataset.max_depth), protocol="test"))
    want = np.concatenate(tabs).mean(0)
    assert abs(res["loss"] - np.mean(losses)) <= 1e-6 * max(1.0, abs(np.mean(losses)))
    for k, name in enumerate(uc.METRIC_NAMES):
        assert abs(res[name] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), name


def test_checkpoint_round_trip_reference_format(tmp_path):
    """train.py:1003-1017 / :600-606: {'epoch','state_dict','optimizer'}; the optimizer entry has torch.optim.AdamW's
    layout, so moments saved by the reference load here and vice versa; training resumes bit-identically."""
    from audio_depth_estimation_b200.checkpoint import checkpoint_path, load_checkpoint, save_checkpoint
    from audio_depth_estimation_b200.config_loader import load_config
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    from audio_depth_estimation_b200.training import TrainStep
    cfg = load_config()
    cfg.dataset.images_size, cfg.model.generator, cfg.model.precision = 128, "unet_128", "fp32"

    def make():
        torch.manual_seed(1)
        net = define_G(cfg, 2, 1, 16, "unet_128", "batch", False, gpu_ids=[0])
        return net, TrainStep(cfg, net, lr=1e-3, waveform_input=False)

    def batch(i):
        return (torch.from_numpy(synthetic.feature_like(2, 128, seed=70 + i)).cuda(),
                torch.from_numpy(synthetic.gt_depth(2, 128, 30.0, seed=80 + i, normalised=False)).cuda())

    net, step = make()
    for i in range(3):
        step(*batch(i))
    path = save_checkpoint(checkpoint_path("exp", 7, root=str(tmp_path)), 7, net, step.optimizer, data_parallel_keys=True)
    assert path.endswith(os.path.join("exp", "checkpoint_7.pth"))
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "state_dict", "optimizer"} and all(k.startswith("module.") for k in ck["state_dict"])
    # a stock torch AdamW over a same-shaped module accepts the optimizer entry (the reference's optimizer class)
    shadow = [torch.nn.Parameter(torch.zeros_like(p, device="cpu")) for p in net.parameters()]
    topt = torch.optim.AdamW(shadow, lr=1e-3)
    topt.load_state_dict({"state": ck["optimizer"]["state"], "param_groups": ck["optimizer"]["param_groups"]})
    assert int(topt.state[shadow[0]]["step"]) == 3 and topt.state[shadow[0]]["exp_avg"].shape == shadow[0].shape
    loss_a = step(*batch(3))
    net2, step2 = make()
    with torch.no_grad():
        for p in net2.parameters():
            p.add_(1.0)                                   # make sure the load is what restores them
    assert load_checkpoint(path, net2, step2.optimizer) == 8
    loss_b = step2(*batch(3))
    assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * abs(float(loss_a))
    for (k, a), b in zip(net.state_dict().items(), net2.state_dict().values()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), k
    # and the reverse direction: moments written by torch.optim.AdamW.state_dict()
    step2.optimizer.load_state_dict(topt.state_dict())
    assert step2.optimizer.step_count == 3


def test_depth_prepare_bit_exact_with_numpy_cv2():
    cv2 = pytest.importorskip("cv2")
    from audio_depth_estimation_b200.feature import DepthTransform
    rng = np.random.default_rng(8)
    for (H, W, S) in ((720, 1280, 256), ((90, 160, 64)), ((100, 100, 256)), ((257, 33, 17))):
        raw = rng.uniform(-800, 45000, size=(3, H, W)).astype(np.float32)
        # V2 (:68-78)
        got = DepthTransform(S, 30.0)(torch.from_numpy(raw).cuda()).cpu().numpy()
        for b in range(3):
            d = raw[b].copy() / 1000.0
            d[d > 30.0] = 30.0
            d[d < 0] = 0
            assert np.array_equal(got[b, 0], cv2.resize(d, (S, S), interpolation=cv2.INTER_NEAREST))
        # V1 (:47-65): nan / inf handling and normalisation
        raw[:, 0, :7] = np.nan; raw[:, 1, :7] = np.inf; raw[:, 2, :7] = -np.inf
        got = DepthTransform(S, 12.0, depth_norm=True, nan_to_num=True)(torch.from_numpy(raw).cuda()).cpu().numpy()
        for b in range(3):
            d = np.nan_to_num(raw[b].copy())
            d = d / 1000
            d[d > 12.0] = 12.0
            d[d < 0.0] = 0.0
            d = cv2.resize(d, (S, S), interpolation=cv2.INTER_NEAREST) / 12.0
            assert np.array_equal(got[b, 0], d.astype(np.float32))
    u16 = rng.integers(0, 65535, size=(2, 96, 128), dtype=np.uint16)
    got = DepthTransform(64, 30.0)(torch.from_numpy(u16.view(np.int16)).cuda()).cpu().numpy()
    for b in range(2):
        d = u16[b].astype(np.float32) / 1000.0
        d[d > 30.0] = 30.0
        assert np.array_equal(got[b, 0], cv2.resize(d, (64, 64), interpolation=cv2.INTER_NEAREST))


def test_bf16_weight_mirror_follows_every_parameter_change():
    """FusedClipAdamW keeps a bf16 copy of the weights for the forward pass; any other writer (load_state_dict,
    in-place edits, a recorded CUDA graph whose weights are replaced) must invalidate it."""
    from audio_depth_estimation_b200.config_loader import load_config
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    from audio_depth_estimation_b200.training import TrainStep
    cfg = load_config()
    cfg.dataset.images_size, cfg.model.generator, cfg.model.precision = 128, "unet_128", "bf16"

    def make(seed):
        torch.manual_seed(seed)
        return define_G(cfg, 2, 1, 16, "unet_128", "batch", False, gpu_ids=[0])

    def batch(i):
        return (torch.from_numpy(synthetic.feature_like(2, 128, seed=170 + i)).cuda(),
                torch.from_numpy(synthetic.gt_depth(2, 128, 30.0, seed=180 + i, normalised=False)).cuda())

    def eval_out(net, x):
        net.eval()
        with torch.no_grad():
            return net(x).clone()

    def close(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-12)) <= 2e-2

    x = batch(9)[0]
    for graph in (False, True):
        net = make(1)
        step = TrainStep(cfg, net, lr=1e-3, waveform_input=False, cuda_graph=graph)
        for i in range(4):
            step(*batch(i))
        assert net._mirror_ok()                                   # the optimiser's mirror is what the forward uses
        trained = eval_out(net, x)
        other = make(2)
        sd = {k: v.clone() for k, v in other.state_dict().items()}
        want = eval_out(other.cuda(), x)
        net.load_state_dict(sd)
        assert not net._mirror_ok()
        got = eval_out(net, x)
        assert close(got, want) and not close(got, trained)
        # training continues from the loaded weights (graph mode records again)
        ref = TrainStep(cfg, other, lr=1e-3, waveform_input=False)
        ref.optimizer.load_state_dict(step.optimizer.state_dict())          # same moments and step count
        for i in range(4, 7):
            la, lb = float(step(*batch(i))), float(ref(*batch(i)))
            assert abs(la - lb) <= (2e-3 if i == 4 else 5e-2) * abs(lb), (graph, i, la, lb)
        with torch.no_grad():
            next(net.parameters()).mul_(0.5)                      # in-place edit bumps the version counter
        assert not net._mirror_ok()
