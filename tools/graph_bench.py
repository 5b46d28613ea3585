"""Eager launches (4 kernels with programmatic dependent launch) vs a CUDA-graph replay of the same step, per batch size."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from squeezedet_pytorch_b200 import ops, synth
dev = torch.device("cuda", 0)
shp = synth.KITTI
w, b = synth.convdet_params(shp, 22)
dw, db = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
a32 = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
packed = ops.pack_convdet_weights(dw)
args = (dw, db, a32, 9, 3, shp.input_hw, shp.top_k, shp.nms_thresh, shp.score_thresh)
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for B in (1, 2, 4, 8, 20):
    feat = torch.from_numpy(synth.features(shp, B, 3)).to(dev)
    out = ops._alloc_detections(B, shp.top_k, dev)
    eager = lambda: ops.head_detect(feat, *args, packed=packed, out=out)
    eager(); torch.cuda.synchronize()
    t_eager = timeit(eager)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.head_detect(feat, *args, packed=packed, out=out)
    t_graph = timeit(g.replay)
    print("B=%-3d eager %7.1f us/step   graph replay %7.1f us/step" % (B, t_eager, t_graph))
