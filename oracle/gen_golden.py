"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from
/root/reference in the build container (CPU, fp32).

Test infrastructure only.  Run:  python oracle/gen_golden.py
/root/reference does not exist on the GPU box, so the vectors are committed and
this script is the provenance record.  Inputs are regenerated from numpy seeds
(audio_depth_estimation_b200/synthetic.py, oracle.unet_oracle.make_state_dict),
only outputs are stored.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from audio_depth_estimation_b200 import synthetic  # noqa: E402
from oracle import unet_oracle  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def ref_cfg(depth_norm, max_depth, size=256):
    return SimpleNamespace(dataset=SimpleNamespace(
        depth_norm=depth_norm, preprocess="resize", images_size=size,
        max_depth=max_depth, audio_format="spectrogram", dataset_dir="/nonexistent"))


def gen_feature():
    import torchaudio.transforms as T  # noqa: F401  (the reference imports it the same way)
    from dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
    from dataloader.BatvisionV1_Dataset import BatvisionV1Dataset
    from dataloader.utils_dataset import get_transform

    out = {}
    # --- V2 chain, full length (BatvisionV2_Dataset.py:96-135) -------------
    cfg = ref_cfg(False, 30.0)
    ds2 = BatvisionV2Dataset.__new__(BatvisionV2Dataset)   # no csv tree on disk
    ds2.cfg = cfg
    for name, echo in (("v2", False), ("v2echo", True)):
        w = torch.from_numpy(synthetic.waveform(1, 8000, seed=11, echo=echo)[0])
        cut = int((2 * cfg.dataset.max_depth / 340) * 44100)
        wc = w[:, :cut]
        spec = ds2._get_spectrogram(wc, n_fft=512, power=1.0, win_length=64, hop_length=16)
        out[name + "_spec_slice"] = spec[:, :, ::37].numpy().copy()
        spec = torch.log(spec + 1e-8)
        for c in range(spec.shape[0]):
            lo, hi = spec[c].min(), spec[c].max()
            spec[c] = (spec[c] - lo) / (hi - lo) if hi > lo else torch.zeros_like(spec[c])
        out[name + "_feat"] = get_transform(cfg, convert=False)(spec).numpy()
    # --- V1 chain (BatvisionV1_Dataset.py:70-78) ---------------------------
    cfg1 = ref_cfg(True, 12.0)
    ds1 = BatvisionV1Dataset.__new__(BatvisionV1Dataset)
    ds1.cfg = cfg1
    w = torch.from_numpy(synthetic.waveform(1, synthetic.V1_LEN, seed=12)[0])
    spec = ds1._get_spectrogram(w, n_fft=512, power=1.0, win_length=64, hop_length=64 // 4)
    out["v1_spec_slice"] = spec[:, :, ::13].numpy().copy()
    out["v1_feat"] = get_transform(cfg1, convert=False)(spec).numpy()
    # --- small full spectrograms, both STFT parameter sets -----------------
    w = torch.from_numpy(synthetic.waveform(1, 1000, seed=13)[0])
    out["small_spec_512"] = ds2._get_spectrogram(w, n_fft=512, power=1.0, win_length=64, hop_length=16).numpy()
    w = torch.from_numpy(synthetic.waveform(1, 2000, seed=14)[0])
    out["small_spec_400"] = ds2._get_spectrogram(w, n_fft=400, power=1.0, win_length=200, hop_length=100).numpy()
    # resize of a non-square ramp+noise plane to a non-256 size
    cfg64 = ref_cfg(False, 30.0, size=64)
    rng = np.random.default_rng(15)
    plane = rng.uniform(0, 1, size=(2, 257, 101)).astype(np.float32)
    out["resize_257x101_to_64"] = get_transform(cfg64, convert=False)(torch.from_numpy(plane)).numpy()
    np.savez_compressed(os.path.join(OUT, "feature.npz"), **out)
    print("feature.npz", {k: v.shape for k, v in out.items()})


def run_ref_step(netG, ngf, batch, size, depth_norm, max_depth, seed, train=True, warm_stats=False):
    from models.unetbaseline_model import define_G
    from utils_loss import SIlogLoss
    num_downs = 8 if netG == "unet_256" else 7
    cfg = ref_cfg(depth_norm, max_depth, size)
    torch.manual_seed(0)
    net = define_G(cfg, 2, 1, ngf, netG, "batch", False, gpu_ids=[])
    sd = unet_oracle.make_state_dict(ngf, num_downs, seed=seed)
    if warm_stats:
        rng = np.random.default_rng(seed + 1)
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = torch.from_numpy(rng.normal(0, 0.05, sd[k].shape).astype(np.float32))
            if k.endswith("running_var"):
                sd[k] = torch.from_numpy(rng.uniform(0.5, 1.5, sd[k].shape).astype(np.float32))
    assert list(net.state_dict().keys()) == list(unet_oracle.ordered_state_dict(sd, num_downs).keys())
    net.load_state_dict(sd, strict=True)
    net.train(train)
    x = torch.from_numpy(synthetic.feature_like(batch, size, seed=seed + 2))
    gt = torch.from_numpy(synthetic.gt_depth(batch, size, max_depth, seed=seed + 3, normalised=depth_norm))
    out = {}
    if not train:
        with torch.no_grad():
            out["y"] = net(x).numpy()
        return out
    # step body, train.py:633-693 (Combined criterion, conf/mode/train.yaml:12-15)
    l1c, sic = torch.nn.L1Loss(), SIlogLoss(lambda_scale=0.869)
    y = net(x)
    mask = gt != 0.0
    if depth_norm:
        p, g = y[mask] * max_depth, gt[mask] * max_depth
    else:
        p, g = y[mask], gt[mask]
    l1, si = l1c(p, g), sic(p, g)
    loss = 0.237 * l1 + 0.637 * si
    y.retain_grad()
    loss.backward()
    out["y"] = y.detach().numpy()
    out["dy"] = y.grad.numpy()
    out["loss"] = np.array([loss.item(), l1.item(), si.item()], dtype=np.float64)
    names = [n for n, _ in net.named_parameters()]
    out["param_names"] = np.array(names)
    out["grad_norms"] = np.array([p.grad.double().norm().item() for _, p in net.named_parameters()])
    out["grad_head"] = np.stack([p.grad.reshape(-1)[:16].numpy() if p.numel() >= 16 else
                                 np.pad(p.grad.reshape(-1).numpy(), (0, 16 - p.numel()))
                                 for _, p in net.named_parameters()])
    sdo = net.state_dict()
    stats = [k for k in sdo if k.endswith("running_mean") or k.endswith("running_var")]
    out["stat_names"] = np.array(stats)
    out["stat_head"] = np.stack([sdo[k][:8].numpy() for k in stats])
    # one optimiser step: clip_grad_norm_(1.0) + AdamW(lr) (train.py:471-476, :689-691)
    opt = torch.optim.AdamW(net.parameters(), lr=0.002)
    tn = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
    opt.step()
    out["total_norm"] = np.array([tn.item()])
    out["param_head_after"] = np.stack([p.detach().reshape(-1)[:16].numpy() if p.numel() >= 16 else
                                        np.pad(p.detach().reshape(-1).numpy(), (0, 16 - p.numel()))
                                        for _, p in net.named_parameters()])
    return out


def gen_unet():
    cases = {
        # name: (netG, ngf, batch, size, depth_norm, max_depth, seed, train, warm)
        "u256_ngf64_b2_relu": ("unet_256", 64, 2, 256, False, 30.0, 100, True, False),
        "u128_ngf16_b3_sigmoid": ("unet_128", 16, 3, 128, True, 12.0, 200, True, False),
        "u128_ngf64_b2_relu": ("unet_128", 64, 2, 128, False, 30.0, 300, True, False),
        "u128_ngf16_b2_eval": ("unet_128", 16, 2, 128, False, 30.0, 400, False, True),
        "u256_ngf64_b1_eval": ("unet_256", 64, 1, 256, True, 12.0, 500, False, True),
    }
    for name, c in cases.items():
        out = run_ref_step(*c)
        out["case"] = np.array([str(c)])
        np.savez_compressed(os.path.join(OUT, "unet_%s.npz" % name), **out)
        print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def gen_loss():
    from utils_loss import SIlogLoss
    out = {}
    for i, (shape, dn, md) in enumerate([((2, 1, 64, 64), False, 30.0), ((3, 1, 32, 32), True, 12.0)]):
        rng = np.random.default_rng(700 + i)
        gt = synthetic.gt_depth(shape[0], shape[2], md, seed=710 + i, normalised=dn)
        pred = (gt + rng.normal(0, 0.1 * (1.0 if dn else md), gt.shape)).astype(np.float32)
        pred[rng.uniform(size=pred.shape) < 0.1] = 0.0      # ReLU-head zeros (< eps)
        p = torch.from_numpy(pred).requires_grad_(True)
        g = torch.from_numpy(gt)
        mask = g != 0.0
        s = md if dn else 1.0
        l1 = torch.nn.L1Loss()(p[mask] * s, g[mask] * s)
        si = SIlogLoss(lambda_scale=0.869)(p[mask] * s, g[mask] * s)
        loss = 0.237 * l1 + 0.637 * si
        loss.backward()
        out["case%d_loss" % i] = np.array([loss.item(), l1.item(), si.item()])
        out["case%d_grad" % i] = p.grad.numpy()
        out["case%d_pred" % i] = pred
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)
    print("loss.npz")


def gen_feature_mel():
    """Mel branch of the V2 transform: BatvisionV2_Dataset.py:111-135, :187-197, the reference's own methods."""
    import torchaudio.transforms as T  # noqa: F401
    from dataloader.BatvisionV2_Dataset import BatvisionV2Dataset
    from dataloader.utils_dataset import get_transform
    out = {}
    cfg = ref_cfg(False, 30.0)
    ds2 = BatvisionV2Dataset.__new__(BatvisionV2Dataset)
    ds2.cfg = cfg
    for name, echo, seed in (("mel", False, 21), ("melecho", True, 22)):
        w = torch.from_numpy(synthetic.waveform(1, 8000, seed=seed, echo=echo)[0])
        wc = w[:, :int((2 * cfg.dataset.max_depth / 340) * 44100)]
        spec = ds2._get_melspectrogram(wc, n_fft=512, power=1.0, win_length=64)
        out[name + "_spec"] = spec.numpy().copy()
        spec = torch.log(spec + 1e-8)
        for c in range(spec.shape[0]):
            lo, hi = spec[c].min(), spec[c].max()
            spec[c] = (spec[c] - lo) / (hi - lo) if hi > lo else torch.zeros_like(spec[c])
        out[name + "_feat"] = get_transform(cfg, convert=False)(spec).numpy()
    w = torch.from_numpy(synthetic.waveform(1, 3000, seed=23)[0])
    out["small_mel_400"] = ds2._get_melspectrogram(w).numpy()             # the method's defaults: n_fft 400, win 400
    np.savez_compressed(os.path.join(OUT, "feature_mel.npz"), **out)
    print("feature_mel.npz", {k: v.shape for k, v in out.items()})


def gen_metrics():
    """utils_criterion.compute_errors after the per-sample preparation of train.py:807-825."""
    from utils_criterion import compute_errors
    out = {}
    cases = [("m30", False, 30.0, 64, 0), ("n12", True, 12.0, 48, 1), ("zeros", False, 30.0, 32, 2),
             ("negpred", False, 30.0, 32, 3), ("emptygt", False, 30.0, 32, 4)]
    for name, dn, md, size, k in cases:
        rng = np.random.default_rng(900 + k)
        gt = synthetic.gt_depth(3, size, md, seed=910 + k, normalised=dn)
        pred = (gt + rng.normal(0, 0.15 * (1.0 if dn else md), gt.shape)).astype(np.float32)
        if name == "zeros":
            pred[:] = 0.0
        if name == "negpred":
            pred = -np.abs(pred) - 1.0
        if name == "emptygt":
            gt[1] = 0.0
        rows = []
        for i in range(gt.shape[0]):
            g, p = gt[i, 0].copy(), pred[i, 0].copy()
            if dn:
                g, p = g * md, p * md
            eps = 1e-3 if dn else 1e-6
            p = np.clip(p, eps, md)
            g = np.maximum(g, 0.0)
            rows.append([float(v) for v in compute_errors(g, p, min_depth_threshold=0.0)])
        out[name + "_pred"] = pred
        out[name + "_errors"] = np.array(rows, dtype=np.float64)
    # compute_errors called directly (no clipping): its own fall-back branches (:38-54)
    rng = np.random.default_rng(950)
    g = rng.uniform(0.5, 20.0, size=(40, 40)).astype(np.float32)
    g[rng.uniform(size=g.shape) < 0.2] = 0.0
    p = (g + rng.normal(0, 2.0, g.shape)).astype(np.float32)
    out["raw_gt"], out["raw_pred"] = g, p
    out["raw_errors"] = np.array([float(v) for v in compute_errors(g, p)])
    out["rawneg_errors"] = np.array([float(v) for v in compute_errors(g, -np.abs(p) - 1.0)])
    out["rawsmall_errors"] = np.array([float(v) for v in compute_errors(g / 100.0, np.abs(p) / 100.0)])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics.npz", {k: v.shape for k, v in out.items()})


def gen_binaural():
    """BASELINE config 4: the unmodified BinauralAttentionDepthNet (models/binaural_attention_model.py:181-340), train-mode
    forward + backward of L = sum(y * r) and an eval-mode forward afterwards.  The weights are the reference's own
    initialisation under torch.manual_seed(0) (the mirror creates its modules in the same order, so the same seed gives
    the same tensors); the attention gammas are set to 0.5 so that the attention branch contributes."""
    from models.binaural_attention_model import BinauralAttentionDepthNet
    out = {}
    for name, levels, batch in (("lv345_b2", [3, 4, 5], 2), ("lv2345_b2", [2, 3, 4, 5], 2)):
        torch.manual_seed(0)
        net = BinauralAttentionDepthNet(base_channels=64, bilinear=True, output_size=128, max_depth=30.0, attention_levels=levels)
        with torch.no_grad():
            for m in net.attention_modules.values():
                m.gamma.fill_(0.5)
            net.outc[0].weight.mul_(0.1)          # un-saturate the sigmoid head of the untrained network (logits ~ N(12, 7))
            net.outc[0].bias.fill_(-1.2)
        x = torch.from_numpy(synthetic.feature_like(batch, 128, seed=301))
        r = torch.from_numpy(np.random.default_rng(302).normal(0, 1, (batch, 1, 128, 128)).astype(np.float32))
        net.train()
        y = net(x)
        (y * r).sum().backward()
        out[name + "_y"] = y.detach().numpy()
        names, norms, heads = [], [], []
        for k, p_ in net.named_parameters():
            g = p_.grad.detach().reshape(-1)
            names.append(k)
            norms.append(float(g.double().norm()))
            heads.append(np.pad(g[:64].numpy(), (0, max(0, 64 - g.numel()))))
        out[name + "_grad_names"] = np.array(names)
        out[name + "_grad_norms"] = np.array(norms)
        out[name + "_grad_heads"] = np.stack(heads)
        out[name + "_rm_fusion3"] = net.fusion_layers["fusion_3"][1].running_mean.numpy().copy()
        net.eval()
        with torch.no_grad():
            out[name + "_y_eval"] = net(x).numpy()
    np.savez_compressed(os.path.join(OUT, "binaural.npz"), **out)
    print("binaural.npz", {k: v.shape for k, v in out.items()})


def gen_binaural_ct():
    """BASELINE config 4, the other decoder: bilinear=False (ConvTranspose2d k2 s2 up-sampling, :65-66) and an output_size
    that differs from the input (the F.interpolate of :322-328).  Same recipe as gen_binaural; weights from the
    reference's own initialisation under torch.manual_seed(0)."""
    from models.binaural_attention_model import BinauralAttentionDepthNet
    out = {}
    torch.manual_seed(0)
    net = BinauralAttentionDepthNet(base_channels=64, bilinear=False, output_size=96, max_depth=30.0, attention_levels=[4, 5])
    with torch.no_grad():
        for m in net.attention_modules.values():
            m.gamma.fill_(0.5)
        net.outc[0].weight.mul_(0.1)
        net.outc[0].bias.fill_(-1.2)
    x = torch.from_numpy(synthetic.feature_like(2, 128, seed=311))
    r = torch.from_numpy(np.random.default_rng(312).normal(0, 1, (2, 1, 96, 96)).astype(np.float32))
    net.train()
    y = net(x)
    (y * r).sum().backward()
    out["y"] = y.detach().numpy()
    names, norms, heads = [], [], []
    for k, p_ in net.named_parameters():
        g = p_.grad.detach().reshape(-1)
        names.append(k)
        norms.append(float(g.double().norm()))
        heads.append(np.pad(g[:64].numpy(), (0, max(0, 64 - g.numel()))))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    out["grad_heads"] = np.stack(heads)
    net.eval()
    with torch.no_grad():
        out["y_eval"] = net(x).numpy()
    np.savez_compressed(os.path.join(OUT, "binaural_ct.npz"), **out)
    print("binaural_ct.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["feature", "feature_mel", "metrics", "loss", "unet", "binaural"]
    for name in which:
        {"feature": gen_feature, "feature_mel": gen_feature_mel, "metrics": gen_metrics, "loss": gen_loss,
         "unet": gen_unet, "binaural": gen_binaural, "binaural_ct": gen_binaural_ct}[name]()
