"""Step time of the fused head (features -> detections) with the candidate scoring in the stand-alone scan kernel
(SQD_FUSED_SCORE=0) or in the GEMM's scorer warps (1).  usage: python tools/fused_score_time.py [batch ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import _lib, ops, synth
dev = torch.device("cuda")
for name, shp, batches in (("kitti", synth.KITTI, [int(x) for x in sys.argv[1:]] or [1, 20, 256]), ("stress", synth.STRESS, [8])):
    w, b = synth.convdet_params(shp, 4321)
    w, b = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
    packed = ops.pack_convdet_weights(w)
    anchors = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
    for B in batches:
        feats = [torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev)) for _ in range(3)]
        dets = {}
        for val in (0, 1, 0, 1):
            with _lib.option("SQD_FUSED_SCORE", val):
                det = ops._alloc_detections(B, shp.top_k, dev)
                step = lambda i: ops.head_detect(feats[i % 3], w, b, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw,
                                                 shp.top_k, shp.nms_thresh, shp.score_thresh, packed=packed, out=det)
                for i in range(5):
                    step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 60 if B <= 32 else 12
                e0.record()
                for i in range(n):
                    step(i)
                e1.record(); torch.cuda.synchronize()
                det.check_status()
                step(0); torch.cuda.synchronize()
                dets[val] = [t.clone() for t in (det.count, det.anchor, det.cls, det.score, det.box)]
                print(f"{name} B={B} fused_score={val}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per step", flush=True)
        same = all(torch.equal(a, c) for a, c in zip(dets[0], dets[1]))
        print(f"   identical detections: {same}", flush=True)
