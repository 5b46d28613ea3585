"""Device-resident throughput of the fused path (features -> detections) for the other BASELINE.json configurations:
    python tools/shape_bench.py kitti 1024      # configs[2]: 2048 images over 2 GPUs -> 1024 per GPU
    python tools/shape_bench.py stress 64       # configs[4]: 512 images of 2496x768 (C = 8, top-256) over 8 GPUs
Not a bench.py line (those configurations are parity-test cases); prints per-stage times from sqd_head_detect_profile."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import ops, synth  # noqa: E402

shp = {"kitti": synth.KITTI, "stress": synth.STRESS}[sys.argv[1]]
B = int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(7)
R = 2
feats = [torch.relu(torch.randn((B, shp.in_channels, *shp.grid_hw), generator=gen, device=dev)) for _ in range(R)]
w, b = synth.convdet_params(shp, 4321)
w, b = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed = ops.pack_convdet_weights(w)
anchors = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
det = ops._alloc_detections(B, shp.top_k, dev)


def step(i):
    return ops.head_detect(feats[i % R], w, b, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw, shp.top_k,
                           shp.nms_thresh, shp.score_thresh, packed=packed, out=det)


for i in range(3):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
rows = []
for i in range(5):
    _, st = ops.head_detect_profile(feats[i % R], b, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw, shp.top_k,
                                    shp.nms_thresh, shp.score_thresh, packed, out=det)
    rows.append(st)
st = np.asarray(rows[1:]).mean(0)
flop = 2 * shp.grid_hw[0] * shp.grid_hw[1] * shp.out_channels * 9 * shp.in_channels
print(f"{shp.name} B={B}: {ms:.3f} ms/step = {B / ms * 1e3:,.0f} images/s | pre-pass {st[0]:.3f} ms, ConvDet {st[1]:.3f} ms "
      f"({B * flop / st[1] / 1e9:.0f} TFLOP/s algorithmic), filter {st[2]:.3f} ms | kept per image: mean {float(det.count.float().mean()):.1f}")
