"""Per-parameter agreement (norm ratio, cosine over the first 64 entries) of the config-4 gradients with the golden file."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from audio_depth_estimation_b200 import synthetic
from audio_depth_estimation_b200.models import binaural_attention_model as bam

name, levels = ("lv345_b2", [3, 4, 5]) if len(sys.argv) < 2 else ("lv2345_b2", [2, 3, 4, 5])
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "binaural.npz"))
torch.manual_seed(0)
net = bam.BinauralAttentionDepthNet(64, True, 128, 30.0, levels)
with torch.no_grad():
    for m in net.attention_modules.values():
        m.gamma.fill_(0.5)
    net.outc[0].weight.mul_(0.1)
    net.outc[0].bias.fill_(-1.2)
net = net.cuda().train()
x = torch.from_numpy(synthetic.feature_like(2, 128, seed=301)).cuda()
r = torch.from_numpy(np.random.default_rng(302).normal(0, 1, (2, 1, 128, 128)).astype(np.float32)).cuda()
y = net(x)
(y * r).sum().backward()
names = list(g[name + "_grad_names"])
params = dict(net.named_parameters())
for i, k in enumerate(names):
    got = params[k].grad.reshape(-1)[:64].double().cpu().numpy()
    ref = g[name + "_grad_heads"][i][:got.size].astype(np.float64)
    cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
    n = float(params[k].grad.double().norm())
    print("%-62s ratio %7.3f  cos %6.3f" % (k, n / max(g[name + "_grad_norms"][i], 1e-30), cos))
