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
