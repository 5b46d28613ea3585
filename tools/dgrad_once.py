import sys, numpy as np, torch
sys.path.insert(0, ".")
from squeezedet_pytorch_b200 import ops, synth
dev = torch.device("cuda")
shp, B = synth.KITTI, 20
w, _ = synth.convdet_params(shp, 2)
w = torch.from_numpy(w).to(dev)
g = torch.randn(B, *shp.grid_hw, shp.out_channels, device=dev)
for _ in range(4): ops.convdet_dgrad(g, w)
torch.cuda.synchronize()
