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

# ---- the inference step -------------------------------------------------------------------------------------------
wf, bf = synth.convdet_params(shp, 2)
wf, bf = torch.from_numpy(wf).to(dev), torch.from_numpy(bf).to(dev)
a32 = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
packed = ops.pack_convdet_weights(wf)
for Bf in (20, 1):
    ff = [torch.from_numpy(synth.features(shp, Bf, 5 + i)).to(dev) for i in range(3 if Bf > 1 else 1)]
    out = ops._alloc_detections(Bf, shp.top_k, dev)
    step = lambda i: ops.head_detect(ff[i % len(ff)], wf, bf, a32, 9, 3, shp.input_hw, shp.top_k, shp.nms_thresh, shp.score_thresh,
                                     packed=packed, out=out)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(30): step(i)
        torch.cuda.synchronize()
    print("-- inference step, B = %d" % Bf)
    ev = [e for e in prof.events() if e.device_time_total > 0]
    t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
    print("   span of 30 steps: %.1f us per step" % ((t1 - t0) / 30))
    for e in sorted(prof.key_averages(), key=lambda r: -r.device_time_total):
        if e.device_time_total > 0:
            print("%9.1f us  x%-3d %s" % (e.device_time_total / e.count, e.count // 30, e.key[:100]))
