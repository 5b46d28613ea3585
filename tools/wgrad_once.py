"""Profiling aid: a few weight-gradient calls (ncu -k regex:wgrad_tc3 ...)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from squeezedet_pytorch_b200 import ops, synth
dev = torch.device("cuda")
shp, B = synth.KITTI, 20
feat = torch.from_numpy(synth.features(shp, B, 1)).to(dev)
g = torch.randn(B, *shp.grid_hw, shp.out_channels, device=dev)
for _ in range(4): ops.convdet_wgrad(feat, g, tensor_cores=True)
torch.cuda.synchronize()
