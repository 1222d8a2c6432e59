"""The dataset mirrors against the reference's behaviour on small BatVision-shaped trees written to tmp_path:
depth preparation (mm -> m, clip, cv2.INTER_NEAREST resize, V1 normalisation / nan handling), file discovery,
'waveform' pass-through (CPU) and the GPU spectrogram branch against the oracle."""
import os
import wave
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest
import torch

from audio_depth_estimation_b200 import synthetic
from audio_depth_estimation_b200.dataloader._common import nearest_resize
from oracle import feature_oracle as fo


def cfg_v2(root, fmt="waveform", size=64):
    return SimpleNamespace(dataset=SimpleNamespace(name="batvisionv2", dataset_dir=str(root), audio_format=fmt, preprocess="resize",
                                                   depth_norm=False, images_size=size, max_depth=30.0))


def cfg_v1(root, fmt="waveform", size=64):
    return SimpleNamespace(dataset=SimpleNamespace(name="batvisionv1", dataset_dir=str(root), audio_format=fmt, preprocess="resize",
                                                   depth_norm=True, images_size=size, max_depth=12.0))


def write_wav(path, x):
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as f:
        f.setnchannels(x.shape[0]); f.setsampwidth(2); f.setframerate(44100)
        f.writeframes(pcm.tobytes())
    return pcm.T.astype(np.float32) / 32768.0


@pytest.fixture()
def v2_tree(tmp_path):
    rng = np.random.default_rng(0)
    waves = {}
    for loc in ("office_a", "hall_b"):
        d = tmp_path / loc
        (d / "depth").mkdir(parents=True); (d / "audio").mkdir()
        rows = []
        for i in range(2):
            depth = rng.uniform(-500, 45000, size=(90, 160)).astype(np.float32)      # mm, some < 0 and > max_depth
            np.save(d / "depth" / ("d%d.npy" % i), depth)
            w = synthetic.waveform(1, 9000, seed=hash((loc, i)) % 1000)[0] * 0.9
            waves[(loc, i)] = write_wav(d / "audio" / ("a%d.wav" % i), w)
            rows.append({"depth path": loc + "/depth", "depth file name": "d%d.npy" % i,         # paths are relative to dataset_dir
                         "audio path": loc + "/audio", "audio file name": "a%d.wav" % i})
        pd.DataFrame(rows).to_csv(d / "train.csv", index=False)
    (tmp_path / "__pycache__").mkdir()
    (tmp_path / "junk_unzipped").mkdir()
    return tmp_path, waves


def test_nearest_resize_is_cv2_inter_nearest():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for shape, size in (((90, 160), 64), ((720, 1280), 256), ((100, 100), 256), ((257, 33), 17)):
        d = rng.uniform(0, 30, size=shape).astype(np.float32)
        assert np.array_equal(nearest_resize(d, size), cv2.resize(d, (size, size), interpolation=cv2.INTER_NEAREST))


def test_v2_dataset_waveform_and_depth(v2_tree):
    cv2 = pytest.importorskip("cv2")
    from audio_depth_estimation_b200.dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
    root, waves = v2_tree
    ds = BatvisionV2Dataset(cfg_v2(root), "train.csv")
    assert len(ds) == 4 and list(ds.instances.columns)[:2] == ["depth path", "depth file name"]
    assert len(BatvisionV2Dataset(cfg_v2(root), "train.csv", location_blacklist=["hall_b"])) == 2
    with pytest.raises(ValueError):
        BatvisionV2Dataset(cfg_v2(root), "missing.csv")
    w, gt = ds[0]
    cut = int((2 * 30.0 / 340) * 44100)
    assert w.shape == (2, cut) and w.dtype == torch.float32
    # depth: the reference's steps (BatvisionV2_Dataset.py:65-78) with cv2 itself
    import os
    first_loc = [d for d in os.listdir(root) if (root / d).is_dir()][0]      # os.listdir order, as the reference (:22-26)
    d = np.load(root / first_loc / "depth" / "d0.npy").astype(np.float32) / 1000.0
    d[d > 30.0] = 30.0
    d[d < 0] = 0
    ref = cv2.resize(d, (64, 64), interpolation=cv2.INTER_NEAREST)
    assert gt.shape == (1, 64, 64) and np.array_equal(gt[0].numpy(), ref)
    assert np.array_equal(w.numpy(), waves[(first_loc, 0)][:, :cut])


def test_v1_dataset_depth_nan_inf_and_blacklist(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from audio_depth_estimation_b200.dataloader.BatvisionV1_Dataset import BatvisionV1Dataset
    rng = np.random.default_rng(2)
    rows = []
    for i, loc in enumerate(("lab", "corridor")):
        depth = rng.uniform(0, 20000, size=(80, 120)).astype(np.float32)
        depth[0, :5] = np.nan; depth[1, :5] = np.inf; depth[2, :5] = -np.inf
        np.save(tmp_path / ("depth_%d.npy" % i), depth)
        for ear in ("left", "right"):
            np.save(tmp_path / ("%s_%s_%d.npy" % (loc, ear, i)), synthetic.waveform(1, synthetic.V1_LEN, seed=10 * i + (ear == "left"))[0, 0])
        rows.append({"depth path": "depth_%d.npy" % i, "audio path left": "%s_left_%d.npy" % (loc, i),
                     "audio path right": "%s_right_%d.npy" % (loc, i)})
    pd.DataFrame(rows).to_csv(tmp_path / "train.csv", index=False)
    ds = BatvisionV1Dataset(cfg_v1(tmp_path), "train.csv")
    assert len(ds) == 2 and len(BatvisionV1Dataset(cfg_v1(tmp_path), "train.csv", location_blacklist=["corridor"])) == 1
    w, gt = ds[0]
    assert w.shape == (2, synthetic.V1_LEN)
    d = np.nan_to_num(np.load(tmp_path / "depth_0.npy").astype(np.float32))
    d[d == -np.inf] = 0; d[d == np.inf] = 0
    d = d / 1000
    d[d > 12.0] = 12.0; d[d < 0.0] = 0.0
    ref = cv2.resize(d, (64, 64), interpolation=cv2.INTER_NEAREST) / 12.0
    assert np.array_equal(gt[0].numpy(), ref.astype(np.float32))


def test_spectrogram_branch_needs_cuda_or_fails_loudly(v2_tree):
    from audio_depth_estimation_b200.dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
    root, _ = v2_tree
    ds = BatvisionV2Dataset(cfg_v2(root, "spectrogram"), "train.csv")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ds[0]


@pytest.mark.gpu
def test_v2_and_v1_spectrogram_getitem_on_gpu(v2_tree, tmp_path):
    from audio_depth_estimation_b200.dataloader.BatvisionV1_Dataset import BatvisionV1Dataset
    from audio_depth_estimation_b200.dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
    root, waves = v2_tree
    ds = BatvisionV2Dataset(cfg_v2(root, "spectrogram", size=256), "train.csv")
    x, gt = ds[1]
    import os
    first_loc = [d for d in os.listdir(root) if (root / d).is_dir()][0]      # os.listdir order, as the reference (:22-26)
    ref = fo.feature_v2(waves[(first_loc, 1)], 30.0, 256)
    assert x.is_cuda and x.shape == (2, 256, 256) and gt.shape == (1, 256, 256)
    assert np.abs(x.cpu().numpy() - ref).max() <= 5e-4
    spec = ds._get_spectrogram(torch.from_numpy(waves[(first_loc, 1)][:, :4000]).cuda(), n_fft=512, power=1.0, win_length=64, hop_length=16)
    assert spec.shape == (2, 257, 251)
    # the default audio_format of conf/dataset/batvisionv2.yaml
    dm = BatvisionV2Dataset(cfg_v2(root, "mel_spectrogram", size=256), "train.csv")
    xm, _ = dm[1]
    refm = fo.feature_v2_mel(waves[(first_loc, 1)], 30.0, 256)
    assert xm.shape == (2, 256, 256) and np.abs(xm.cpu().numpy() - refm).max() <= 5e-4
    assert dm._get_melspectrogram(torch.from_numpy(waves[(first_loc, 1)][:, :4000]).cuda(), n_fft=512, win_length=64).shape == (2, 32, 126)
    # V1
    v1 = tmp_path / "v1"
    v1.mkdir()
    np.save(v1 / "depth.npy", np.full((60, 60), 5000.0, dtype=np.float32))
    wl, wr = synthetic.waveform(1, synthetic.V1_LEN, seed=77)[0]
    np.save(v1 / "l.npy", wl); np.save(v1 / "r.npy", wr)
    pd.DataFrame([{"depth path": "depth.npy", "audio path left": "l.npy", "audio path right": "r.npy"}]).to_csv(v1 / "train.csv", index=False)
    d1 = BatvisionV1Dataset(cfg_v1(v1, "spectrogram", size=256), "train.csv")
    x1, g1 = d1[0]
    ref1 = fo.feature_v1(np.stack((wl, wr)), 256)
    assert np.abs(x1.cpu().numpy() - ref1).max() <= 1e-4 * np.abs(ref1).max()
    assert np.allclose(g1.numpy(), 5.0 / 12.0)


def test_get_transform_depth_norm_appends_the_reference_minmaxnorm():
    """dataloader/utils_dataset.py:23-27, :31-49: get_transform(depth_norm=True) ends with MinMaxNorm(0, max_depth);
    floats normalise the whole tensor, 2-tuples normalise the two channels of a [2, ...] tensor separately."""
    import torch
    from types import SimpleNamespace
    from audio_depth_estimation_b200.dataloader.utils_dataset import MinMaxNorm, get_transform
    cfg = SimpleNamespace(dataset=SimpleNamespace(preprocess="none", images_size=256, max_depth=12.0))
    t = get_transform(cfg, convert=False, depth_norm=True)
    x = torch.tensor([[[0.0, 3.0], [6.0, 12.0]]])
    assert torch.equal(t(x), x / 12.0)
    assert len(get_transform(cfg, convert=False, depth_norm=False).transforms) == 0
    two = MinMaxNorm(min=(0.0, 1.0), max=(2.0, 5.0))
    y = two(torch.tensor([[1.0, 2.0], [3.0, 5.0]]))
    assert torch.equal(y, torch.tensor([[0.5, 1.0], [0.5, 1.0]]))
    with pytest.raises(AssertionError):
        MinMaxNorm(min=0, max=12)          # (ints are rejected, as in the reference)


def test_v2_dataset_camera_branch(v2_tree):
    """use_image=True (BatvisionV2_Dataset.py:87-90, :199-210): the RGB camera frame, resized and scaled to [0, 1], takes the
    place of the echo; the depth target is the same as on the audio path."""
    cv2 = pytest.importorskip("cv2")
    from audio_depth_estimation_b200.dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
    root, _ = v2_tree
    rng = np.random.default_rng(7)
    for loc in ("office_a", "hall_b"):
        (root / loc / "camera").mkdir()
        df = pd.read_csv(root / loc / "train.csv")
        df["camera path"] = loc + "/camera"
        df["camera file name"] = ["c%d.png" % i for i in range(len(df))]
        for name in df["camera file name"]:
            cv2.imwrite(str(root / loc / "camera" / name), rng.integers(0, 256, size=(72, 128, 3), dtype=np.uint8))
        df.to_csv(root / loc / "train.csv", index=False)
    ds = BatvisionV2Dataset(cfg_v2(root), "train.csv", use_image=True)
    ref = BatvisionV2Dataset(cfg_v2(root), "train.csv")
    img, gt = ds[1]
    inst = ds.instances.iloc[1]
    bgr = cv2.imread(os.path.join(str(root), inst["camera path"], inst["camera file name"]))
    want = cv2.resize(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), (64, 64)).astype(np.float32) / 255.0
    assert img.shape == (3, 64, 64) and img.dtype == torch.float32
    assert np.array_equal(img.numpy(), want.transpose(2, 0, 1))
    assert torch.equal(gt, ref[1][1])
    with pytest.raises(RuntimeError):
        ds._load_image(str(root / "missing.png"))


def test_get_transform_convert_is_totensor():
    """dataloader/utils_dataset.py:14-16: convert=True starts the chain with ToTensor (HWC uint8 -> CHW float in [0, 1];
    float arrays keep their values; a 2-D array gains a channel axis)."""
    from audio_depth_estimation_b200.dataloader.utils_dataset import get_transform
    cfg = SimpleNamespace(dataset=SimpleNamespace(preprocess="none", images_size=64, max_depth=12.0))
    t = get_transform(cfg, convert=True)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(5, 7, 3), dtype=np.uint8)
    out = t(img).cpu()
    assert out.shape == (3, 5, 7) and out.dtype == torch.float32
    assert np.array_equal(out.numpy(), img.transpose(2, 0, 1).astype(np.float32) / 255.0)
    depth = rng.uniform(0, 12, size=(5, 7)).astype(np.float32)
    out = get_transform(cfg, convert=True, depth_norm=True)(depth).cpu()
    assert out.shape == (1, 5, 7) and np.allclose(out.numpy()[0], depth / 12.0, rtol=1e-6)
