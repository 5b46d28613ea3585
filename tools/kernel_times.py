"""Per-kernel durations of the head's backward (dgrad, wgrad, bias grad) from torch.profiler (CUPTI), KITTI B = 20."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from squeezedet_pytorch_b200 import ops, synth
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
shp, B = synth.KITTI, 20
feat = torch.from_numpy(synth.features(shp, B, 1)).to(dev)
w, _ = synth.convdet_params(shp, 2)
w = torch.from_numpy(w).to(dev)
g = torch.randn(B, *shp.grid_hw, shp.out_channels, device=dev)
def run():
    ops.convdet_dgrad(g, w); ops.convdet_wgrad(feat, g, tensor_cores=True); ops.convdet_bias_grad(g)
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): run()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / max(e.count, 1), e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, n in sorted(rows, key=lambda r: -r[1]):
    print("%9.1f us  x%-3d %s" % (t, n // 10 if n >= 10 else n, k[:110]))
