"""How well-conditioned is d(gamma) of the cross-attention residual?  Prints |sum dy*att| / sum |dy*att| per call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from audio_depth_estimation_b200 import synthetic
from audio_depth_estimation_b200.models import binaural_attention_model as bam

orig = bam._Residual.backward


def patched(ctx, dy):
    att, gamma = ctx.saved_tensors
    p = dy.float() * att.float()
    print("residual C=%d: sum %.4g  sum|.| %.4g  ratio %.3g" % (att.shape[-1], float(p.sum()), float(p.abs().sum()),
                                                                 float(p.sum().abs() / p.abs().sum())))
    return orig(ctx, dy)


bam._Residual.backward = staticmethod(patched)
# unit check of the op itself
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(2, 8, 8, 128, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
b = torch.randn(2, 8, 8, 128, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
gm = torch.full((1,), 0.5, device="cuda", requires_grad=True)
y = bam._Residual.apply(a, b, gm)
dy = torch.randn_like(y)
y.backward(dy)
ref = (dy.float() * b.float()).sum()
print("unit: dgamma %.5f ref %.5f" % (float(gm.grad), float(ref)))

torch.manual_seed(0)
net = bam.BinauralAttentionDepthNet(64, True, 128, 30.0, [3, 4, 5])
with torch.no_grad():
    for m in net.attention_modules.values():
        m.gamma.fill_(0.5)
    net.outc[0].weight.mul_(0.1)
    net.outc[0].bias.fill_(-1.2)
net = net.cuda().train()
x = torch.from_numpy(synthetic.feature_like(2, 128, seed=301)).cuda()
r = torch.from_numpy(np.random.default_rng(302).normal(0, 1, (2, 1, 128, 128)).astype(np.float32)).cuda()
(net(x) * r).sum().backward()
for k, m in net.attention_modules.items():
    print(k, float(m.gamma.grad))
