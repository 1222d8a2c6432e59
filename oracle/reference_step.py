"""The training / inference step of the reference, driven through the reference's OWN modules (baseline/_ref/, a
git-ignored copy made by oracle/vendor_reference.py; /root/reference in the build container) on the host CPU, with the
oracle port (oracle/unet_oracle.py, oracle/feature_oracle.py) as the fall-back when neither is present.

Measurement infrastructure only: imported by bench.py's `--impl reference` arm and its `cpu_baseline` leg.

The reference's CLIs cannot be imported (train.py:10 pulls in matplotlib), so the 60-line step body train.py:633-693 is
restated here around the unmodified modules:
  * features: BatvisionV2Dataset._get_spectrogram(512, 64, 16) -> log(x + 1e-8) -> per-channel min-max -> get_transform
    Resize, one sample at a time, exactly the loop of BatvisionV2_Dataset.py:96-135 (__getitem__);
  * model: define_G(cfg, 2, 1, 64, 'unet_256', 'batch', False, gpu_ids=[]) (train.py:381);
  * criterion: valid = gt != 0; 0.237 * L1Loss + 0.637 * SIlogLoss(0.869) (train.py:646-669, conf/mode/train.yaml:12-15);
  * optimiser: clip_grad_norm_(1.0) + torch.optim.AdamW(lr) (train.py:471-476, :689-691).
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for root in (os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(root, "models", "unetbaseline_model.py")):
            return root
    return None


def _import_reference(root):
    """Import the reference packages `models`, `dataloader`, `utils_loss` from `root` without leaving it on sys.path."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("models", "dataloader", "utils_loss") or k.startswith(("models.", "dataloader."))}
    sys.path.insert(0, root)
    try:
        from models.unetbaseline_model import define_G
        from utils_loss import SIlogLoss
        from dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
        from dataloader.utils_dataset import get_transform
    finally:
        sys.path.remove(root)
        for k in list(sys.modules):
            if k in ("models", "dataloader", "utils_loss") or k.startswith(("models.", "dataloader.")):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return define_G, SIlogLoss, BatvisionV2Dataset, get_transform


def make_step(B, threads, lr=0.002, train=True, seed=1234):
    """Returns (step_fn, kind, describe): step_fn() runs one training step (train=True) or one eval-mode forward from
    waveforms (train=False) over a fixed synthetic batch of B samples and returns the loss (or 0.0)."""
    from audio_depth_estimation_b200 import synthetic
    torch.set_num_threads(threads)
    wave = synthetic.waveform(B, synthetic.V2_LEN, seed=seed)
    gt = torch.from_numpy(synthetic.gt_depth(B, 256, 30.0, seed=seed + 3087))
    root = reference_root()
    if root is None:
        return _make_port_step(B, wave, gt, lr, train) + ("port",)
    define_G, SIlogLoss, BatvisionV2Dataset, get_transform = _import_reference(root)
    cfg = SimpleNamespace(dataset=SimpleNamespace(depth_norm=False, preprocess="resize", images_size=256, max_depth=30.0,
                                                  audio_format="spectrogram", dataset_dir="/nonexistent"))
    ds = BatvisionV2Dataset.__new__(BatvisionV2Dataset)        # no csv tree on disk: only the transform methods are used
    ds.cfg = cfg
    resize = get_transform(cfg, convert=False)
    torch.manual_seed(0)
    net = define_G(cfg, 2, 1, 64, "unet_256", "batch", False, gpu_ids=[])
    l1c, sic = torch.nn.L1Loss(), SIlogLoss(lambda_scale=0.869)
    opt = torch.optim.AdamW(net.parameters(), lr=lr)
    wave_t = torch.from_numpy(wave)

    def features():
        out = []
        for b in range(B):                                       # BatvisionV2_Dataset.py:96-135, per sample
            spec = ds._get_spectrogram(wave_t[b], n_fft=512, power=1.0, win_length=64, hop_length=16)
            spec = torch.log(spec + 1e-8)
            for c in range(spec.shape[0]):
                lo, hi = spec[c].min(), spec[c].max()
                spec[c] = (spec[c] - lo) / (hi - lo) if hi > lo else torch.zeros_like(spec[c])
            out.append(resize(spec))
        return torch.stack(out)

    def train_step():
        net.train()
        x = features()
        opt.zero_grad()
        y = net(x)
        mask = gt != 0.0
        loss = 0.237 * l1c(y[mask], gt[mask]) + 0.637 * sic(y[mask], gt[mask])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
        opt.step()
        return float(loss.detach())

    def infer_step():
        net.eval()
        with torch.no_grad():
            net(features())
        return 0.0

    what = "unmodified reference modules (%s): BatvisionV2Dataset._get_spectrogram + get_transform, define_G(unet_256), SIlogLoss, torch AdamW" % \
           ("baseline/_ref" if root.startswith(REPO) else root)
    return (train_step if train else infer_step), what, "reference"


def _make_port_step(B, wave, gt, lr, train):
    from oracle import feature_oracle as fo
    from oracle import unet_oracle as uo
    sd = uo.make_state_dict(64, 8, seed=0)
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))]
    for n in names:
        sd[n].requires_grad_(True)
    m = [torch.zeros_like(sd[n]) for n in names]
    v = [torch.zeros_like(sd[n]) for n in names]
    state = {"step": 0}

    def train_step():
        state["step"] += 1
        x = torch.from_numpy(np.stack([fo.feature_v2(wave[b], 30.0, 256) for b in range(B)]))
        y = uo.unet_forward(x, sd, 8, False, training=True)
        loss = uo.depth_loss(y, gt)
        for n in names:
            sd[n].grad = None
        loss.backward()
        with torch.no_grad():
            uo.clip_adamw_step([sd[n] for n in names], [sd[n].grad for n in names], m, v, state["step"], lr)
        return float(loss.detach())

    def infer_step():
        with torch.no_grad():
            x = torch.from_numpy(np.stack([fo.feature_v2(wave[b], 30.0, 256) for b in range(B)]))
            uo.unet_forward(x, sd, 8, False, training=False)
        return 0.0
    return (train_step if train else infer_step), "oracle port of train.py:633-693 (numpy features, torch fp32 U-Net)"
