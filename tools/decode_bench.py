"""Stand-alone decode + NMS on the HBM-bound shape (same measurement as bench.py's roofline_decode_nms / roofline_decode):
N images of SURVEY 8d's clustered pred set resident in HBM.  usage: python tools/decode_bench.py [N=1024] [reps=12]
Small enough to put under `ncu -k regex:"score_candidates|detect_from_candidates|decode_kernel"`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
shp = synth.KITTI
dev = torch.device("cuda")
a64 = synth.anchor_table(shp)
anchors = torch.from_numpy(a64.astype(np.float32)).to(dev)
base = torch.from_numpy(synth.clustered_pred(shp, 16, 777, anchors=a64)).to(dev)
pred = base.repeat((N + 15) // 16, 1, 1)[:N].contiguous()
det = ops._alloc_detections(N, shp.top_k, dev)
dense = None
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
for i in range(reps + 2):
    ev = evs[max(0, i - 2)]
    ev[0].record()
    ops.detect_from_pred(pred, anchors, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh, out=det)
    ev[1].record()
    dense = ops.decode_scores(pred, anchors, shp.input_hw, shp.num_classes, out=dense)
    ev[2].record()
torch.cuda.synchronize()
t_det = float(np.mean([e[0].elapsed_time(e[1]) for e in evs])) * 1e-3
t_dec = float(np.mean([e[1].elapsed_time(e[2]) for e in evs])) * 1e-3
pb, db = 16848 * 8 * 4, 16848 * 28
print(f"{N} images: detect_from_pred {t_det * 1e6:.1f} us = {N * pb / t_det / 1e9:.0f} GB/s ({N / t_det / 1e6:.2f} M img/s) | "
      f"decode_scores {t_dec * 1e6:.1f} us = {N * (pb + db) / t_dec / 1e9:.0f} GB/s | kept/image {float(det.count.float().mean()):.1f}")
