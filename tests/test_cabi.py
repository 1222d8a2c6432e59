"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/adp_b200.h declares; the Python mirror keeps the reference's names and state_dict."""
import os
import re
from types import SimpleNamespace

import pytest
import torch

from audio_depth_estimation_b200 import _lib, build
from oracle import unet_oracle as uo

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(REPO, "include", "adp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adp_[a-z0-9_A-Z]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_error_string(lib):
    assert lib.adp_version() >= 100
    assert isinstance(lib.adp_last_error(), bytes)


def test_argument_errors_do_not_touch_the_gpu(lib):
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.adp_stft_mag(None, 0, 0, 0, 512, 64, 16, None, None) != 0
    assert b"stft" in lib.adp_last_error()
    d = _lib.UnetDesc()
    assert lib.adp_unet_workspace_bytes(d) == 0
    d.batch, d.in_ch, d.out_ch, d.ngf, d.num_downs, d.size, d.dtype = 2, 2, 1, 64, 8, 256, _lib.ADP_BF16
    assert lib.adp_unet_workspace_bytes(d) > 0


def cfg(depth_norm=False):
    return SimpleNamespace(dataset=SimpleNamespace(depth_norm=depth_norm, max_depth=30.0, images_size=256,
                                                   preprocess="resize", name="batvisionv2"),
                           model=SimpleNamespace(precision="fp32"))


@pytest.mark.parametrize("netG,nd,ngf", [("unet_256", 8, 64), ("unet_128", 7, 16)])
def test_state_dict_layout_matches_reference(netG, nd, ngf):
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    net = define_G(cfg(), 2, 1, ngf, netG, "batch", False, gpu_ids=[])
    ref = uo.ordered_state_dict(uo.make_state_dict(ngf, nd, seed=1), nd)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    net.load_state_dict(ref, strict=True)
    net.load_state_dict({"module." + k: v for k, v in ref.items()}, strict=True)   # DataParallel checkpoints
    for k in ref:
        assert torch.equal(net.state_dict()[k], ref[k]), k
    if netG == "unet_256" and ngf == 64:
        assert sum(p.numel() for p in net.parameters()) == 54408833


def test_product_path_refuses_cpu_tensors():
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    from audio_depth_estimation_b200.utils_loss import SIlogLoss
    from audio_depth_estimation_b200 import feature
    net = define_G(cfg(), 2, 1, 16, "unet_128", "batch", False, gpu_ids=[])
    with pytest.raises(_lib.AdpError):
        net(torch.zeros(1, 2, 128, 128))
    with pytest.raises(_lib.AdpError):
        SIlogLoss()(torch.ones(4), torch.ones(4))
    with pytest.raises(_lib.AdpError):
        feature.spectrogram(torch.zeros(2, 1000), 512, 1.0, 64, 16)


def test_unknown_generator_and_norm_raise_like_the_reference():
    from audio_depth_estimation_b200.models.unetbaseline_model import define_G
    with pytest.raises(NotImplementedError):
        define_G(cfg(), 2, 1, 64, "resnet_9blocks")
    with pytest.raises(NotImplementedError):
        define_G(cfg(), 2, 1, 64, "unet_256", norm="group")


def test_config_loader_keys():
    from audio_depth_estimation_b200.config_loader import load_config
    c = load_config("batvisionv2", "train", "exp")
    assert c.mode.experiment_name == "exp" and c.mode.criterion == "Combined"
    assert (c.mode.l1_weight, c.mode.silog_weight, c.mode.silog_lambda) == (0.237, 0.637, 0.869)
    assert c.dataset.max_depth == 30.0 and c.dataset.images_size == 256 and c.dataset.depth_norm is False
    c1 = load_config("batvisionv1", "test")
    assert c1.dataset.depth_norm is True and c1.dataset.max_depth == 12.0 and c1.mode.batch_size == 1
    assert c1.mode.stat_dir == "./eval/" and c1.mode.num_threads == 4 and c1.mode.eval_on == "test"     # conf/mode/test.yaml as shipped by the reference
    assert c.model.generator == "unet_256"


def test_binaural_mirror_state_dict_and_init_match_reference_layout(golden_dir):
    """config 4: same state_dict keys/shapes as models/binaural_attention_model.py (names recorded in the golden file by the
    unmodified reference) and the same parameter order."""
    import numpy as np
    import torch
    from audio_depth_estimation_b200.models.binaural_attention_model import BinauralAttentionDepthNet
    g = np.load(os.path.join(golden_dir, "binaural.npz"))
    torch.manual_seed(0)
    net = BinauralAttentionDepthNet(64, True, 128, 30.0, [2, 3, 4, 5])
    names = [k for k, _ in net.named_parameters()]
    assert names == list(g["lv2345_b2_grad_names"])
    assert net.get_num_params() == 29260773                 # 29,260,773 parameters at levels 2-5
    sd = net.state_dict()
    assert sd["left_encoder.inc.double_conv.0.weight"].shape == (64, 1, 3, 3)
    assert sd["attention_modules.attn_2.query.weight"].shape == (16, 128, 1, 1)
    assert sd["fusion_layers.fusion_5.0.weight"].shape == (512, 1024, 1, 1) and "outc.0.bias" in sd
    net.load_state_dict({"module." + k: v for k, v in sd.items()})
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 2, 128, 128))                     # no CPU fallback


def test_new_host_entry_points_refuse_cpu_and_unsupported_configs(tmp_path):
    """No CPU fallbacks anywhere on the product path, and the config-4 mirror states its limits instead of degrading."""
    import torch
    from audio_depth_estimation_b200 import feature, utils_criterion
    from audio_depth_estimation_b200.checkpoint import checkpoint_path
    from audio_depth_estimation_b200.models.binaural_attention_model import BinauralAttentionDepthNet, Up
    from audio_depth_estimation_b200.optim import FusedClipAdamWParams
    with pytest.raises(RuntimeError):
        feature.DepthTransform(64, 30.0)(torch.zeros(1, 8, 8))
    with pytest.raises(RuntimeError):
        feature.melspectrogram(torch.zeros(2, 4000), n_fft=512, win_length=64)
    with pytest.raises(RuntimeError):
        utils_criterion.batch_errors(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 4, 4), depth_norm=False, max_depth=30.0)
    with pytest.raises(ValueError):
        utils_criterion.batch_errors(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 4, 4), depth_norm=False, max_depth=30.0, protocol="x")
    with pytest.raises(RuntimeError):
        FusedClipAdamWParams([torch.nn.Parameter(torch.zeros(4))])
    with pytest.raises(NotImplementedError):
        BinauralAttentionDepthNet(base_channels=32)
    up = Up(128, 64, bilinear=False)          # round 2: the transposed-conv decoder is built (reference :65-67)
    assert isinstance(up.up, torch.nn.ConvTranspose2d) and up.up.weight.shape == (128, 64, 2, 2)
    assert up.conv.double_conv[0].weight.shape == (64, 128, 3, 3)
    sd_ct = BinauralAttentionDepthNet(64, False, 96, 30.0, [5]).state_dict()
    assert sd_ct["up1.up.weight"].shape == (1024, 512, 2, 2) and sd_ct["left_encoder.down4.maxpool_conv.1.double_conv.3.weight"].shape[0] == 1024
    assert checkpoint_path("exp", 3, root=str(tmp_path)).endswith(os.path.join("exp", "checkpoint_3.pth"))
    # every conv weight of the mirror lives in channels_last memory (the layout the kernels read), also after .to()/.float()
    net = BinauralAttentionDepthNet(64, True, 128, 30.0, [5]).float()
    w = net.up1.conv.double_conv[0].weight
    assert w.shape == (512, 1024, 3, 3) and w.stride() == (9 * 1024, 1, 3 * 1024, 1024)
