import sys; sys.path.insert(0,'.')
from types import SimpleNamespace
import numpy as np, torch
from audio_depth_estimation_b200 import synthetic
from audio_depth_estimation_b200.models.unetbaseline_model import define_G
from oracle import unet_oracle as uo
for (B,S,ngf,nd,netG) in [(4,128,64,7,"unet_128"),(16,128,64,7,"unet_128"),(8,256,64,8,"unet_256")]:
    sd = uo.make_state_dict(ngf, nd, seed=7)
    x = synthetic.feature_like(B, S, seed=3)
    with torch.no_grad():
        ref = uo.unet_forward(torch.from_numpy(x), {k:v.clone() for k,v in sd.items()}, nd, False, training=True)
    for prec in ("fp32","bf16"):
        cfg = SimpleNamespace(dataset=SimpleNamespace(depth_norm=False,max_depth=30.0,images_size=S,preprocess="resize",name="batvisionv2"),model=SimpleNamespace(precision=prec))
        net = define_G(cfg,2,1,ngf,netG,"batch",False,gpu_ids=[0]); net.load_state_dict(uo.ordered_state_dict({k:v.clone() for k,v in sd.items()},nd)); net.train()
        with torch.no_grad(): y = net(torch.from_numpy(x).cuda()).cpu()
        print(B,S,prec,"relL2 %.3e  max/max %.3e  mean|err|/mean|ref| %.3e"%(float((y-ref).norm()/ref.norm()), float((y-ref).abs().max()/ref.abs().max()), float((y-ref).abs().mean()/ref.abs().mean())))
