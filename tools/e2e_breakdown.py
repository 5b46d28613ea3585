"""Where the e2e step's time goes: raw copy+sync wall time vs head_detect_host at several chunk sizes."""
import sys, time, torch
sys.path.insert(0, ".")
from squeezedet_pytorch_b200 import ops, synth
dev = torch.device("cuda", 0)
shp = synth.KITTI
B = 20
feats = torch.from_numpy(synth.features(shp, B, 1)).to(dev) if hasattr(synth, "features") else torch.randn(B, shp.in_channels, *shp.grid_hw, device=dev).relu()
weight, bias = synth.convdet_params(shp, 0) if hasattr(synth, "convdet_params") else (None, None)
if weight is None:
    cout = shp.anchors_per_grid * (shp.num_classes + 5)
    weight = torch.randn(cout, shp.in_channels, 3, 3, device=dev) * 0.002
    bias = torch.zeros(cout, device=dev)
else:
    weight, bias = torch.as_tensor(weight).to(dev), torch.as_tensor(bias).to(dev)
anchors = torch.from_numpy(synth.anchor_table(shp)).float().to(dev)
packed = ops.pack_convdet_weights(weight)
hf = torch.empty(feats.shape, dtype=torch.float32).pin_memory(); hf.copy_(feats)
dst = torch.empty_like(feats)
out = ops.HostDetections(B, shp.top_k)
def wall(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
def copy_only():
    dst.copy_(hf, non_blocking=True); torch.cuda.current_stream().synchronize()
print("copy+sync            %8.1f us" % wall(copy_only))
for ch in (20, 10, 5, 4, 2, 1):
    f = lambda: ops.head_detect_host(hf, weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw, shp.top_k,
                                     shp.nms_thresh, shp.score_thresh, packed=packed, out=out, chunk_images=ch, sync=True)
    print("host call chunk=%-2d   %8.1f us" % (ch, wall(f)))
f = lambda: ops.head_detect_host(hf, weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw, shp.top_k,
                                 shp.nms_thresh, shp.score_thresh, packed=packed, out=out, chunk_images=5, overlap=False, sync=True)
print("host call chunk=5 no overlap %8.1f us" % wall(f))

# ---- serving loop (sync=False, two slots) ----------------------------------------------------------------
outs = [ops.HostDetections(B, shp.top_k) for _ in range(2)]
hfs = [hf, hf.clone().pin_memory()]
def serve(n, ch):
    pend = None
    t_issue = 0.0
    for i in range(n):
        t0 = time.perf_counter()
        d = ops.head_detect_host(hfs[i % 2], weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw, shp.top_k,
                                 shp.nms_thresh, shp.score_thresh, packed=packed, out=outs[i % 2], chunk_images=ch, sync=False, slot=i % 2)
        t_issue += time.perf_counter() - t0
        if pend is not None: pend.wait()
        pend = d
    pend.wait()
    return t_issue / n * 1e6
for ch in (20, 10, 5, 2):
    serve(3, ch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ti = serve(30, ch)
    print("serving loop chunk=%-2d  %8.1f us/step   (host issue %6.1f us/call)" % (ch, (time.perf_counter() - t0) / 30 * 1e6, ti))
# pure copies in the same structure
cs = torch.cuda.Stream()
dsts = [torch.empty_like(feats) for _ in range(2)]
def copies(n):
    evs = [None, None]
    for i in range(n):
        with torch.cuda.stream(cs):
            dsts[i % 2].copy_(hfs[i % 2], non_blocking=True)
            e = torch.cuda.Event(); e.record(cs)
        if evs[(i + 1) % 2] is not None: evs[(i + 1) % 2].synchronize()
        evs[i % 2] = e
    torch.cuda.synchronize()
copies(3)
t0 = time.perf_counter(); copies(30)
print("pure copy loop        %8.1f us/step" % ((time.perf_counter() - t0) / 30 * 1e6))
