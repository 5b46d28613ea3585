"""Phase timeline of detect_from_candidates_kernel (profiling build: SQD_BUILD_TRACE=1 python csrc/build.py, then run with
SQD_LIB_PATH=.../libsqdet_b200_trace.so).  usage: python tools/tail_trace.py [batch] [bench|clustered]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import _lib, ops, synth
dev = torch.device("cuda")
shp = synth.KITTI
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = sys.argv[2] if len(sys.argv) > 2 else "bench"
anchors = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
if kind == "bench":
    feat = torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev))
    w, b = synth.convdet_params(shp, 4321)
    pred = ops.convdet_forward(feat, torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev), num_fields=shp.num_fields)
else:
    base = torch.from_numpy(synth.clustered_pred(shp, 16, 777)).to(dev)
    pred = base.repeat((B + 15) // 16, 1, 1)[:B].contiguous()
names = ["entry", "pdl wait", "keys loaded / selected", "rank sort", "boxes decoded", "IoU masks", "NMS sweep", "emit"]
for img in (0, B // 2, B - 1):
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    os.environ["SQD_TAIL_TRACE"] = hex(trace.data_ptr())
    os.environ["SQD_TAIL_TRACE_IMG"] = str(img)
    for _ in range(3):
        det = ops.detect_from_pred(pred, anchors, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
    torch.cuda.synchronize()
    t = trace.cpu().numpy()
    print(f"image {img}: candidates {t[8]}, sorted {t[9]}, kept {int(det.count[img])}; cycles: " +
          " | ".join(f"{names[i]} +{t[i] - t[i - 1]}" for i in range(1, 8)) + f" | total {t[7] - t[0]}" +
          (f" || select: keys in regs +{t[10] - t[1]} | hist zeroed +{t[11] - t[10]} | hist built +{t[12] - t[11]} | bin found +{t[13] - t[12]} | collected +{t[2] - t[13]}" if t[10] else ""))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.detect_from_pred(pred, anchors, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
e1.record(); torch.cuda.synchronize()
print(f"detect_from_pred (scan + tail), B={B} {kind}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
