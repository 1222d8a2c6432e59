"""Debug probe: loss trajectories of the eager step run twice and of the CUDA-graph step (fp32, tiny batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from audio_depth_estimation_b200 import synthetic
from audio_depth_estimation_b200.models.unetbaseline_model import define_G
from audio_depth_estimation_b200.training import TrainStep
from oracle import unet_oracle as uo
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
cfg = SimpleNamespace(dataset=SimpleNamespace(depth_norm=False, max_depth=30.0, images_size=128, preprocess="resize", name="batvisionv2"),
                      mode=SimpleNamespace(criterion="Combined", l1_weight=0.237, silog_weight=0.637, silog_lambda=0.869, learning_rate=0.002),
                      model=SimpleNamespace(precision=prec))
sd = uo.ordered_state_dict(uo.make_state_dict(16, 7, seed=42), 7)
cuda = lambda a: torch.from_numpy(a).cuda()
batches = [(cuda(synthetic.waveform(2, synthetic.V2_LEN, seed=50 + i)), cuda(synthetic.gt_depth(2, 128, 30.0, seed=60 + i))) for i in range(6)]
for tag, graph in (("eager", False), ("eager", False), ("graph", True), ("graph", True)):
    net = define_G(cfg, 2, 1, 16, "unet_128", "batch", False, gpu_ids=[0])
    net.load_state_dict({k: v.clone() for k, v in sd.items()})
    step = TrainStep(cfg, net, lr=0.002, cuda_graph=graph)
    print(tag, np.array([float(step(w, g)) for w, g in batches]))
    torch.cuda.synchronize()
