"""Quick on-GPU check of the one-kernel ConvDet path against the staged path and float64 (tiny + KITTI), with timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import _lib, ops, synth
from oracle import oracle as orc
dev = torch.device("cuda")
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
for shp, B in ((synth.TINY, 2), (synth.TINY, 5), (synth.KITTI, 1), (synth.KITTI, 3), (synth.KITTI, 20)):
    feat = synth.features(shp, B, 7 + B)
    w, b = synth.convdet_params(shp, 11)
    staged = ops.convdet_forward(d(feat), d(w), d(b), check_status=True)
    with _lib.option("SQD_HEAD_ONE_KERNEL", 1):
        fused = ops.convdet_forward(d(feat), d(w), d(b), check_status=True)
        again = ops.convdet_forward(d(feat), d(w), d(b), check_status=True)
    diff = (fused - staged).abs().max().item()
    print(f"{shp.name} B={B}: max|fused-staged| = {diff:.3e}  deterministic={torch.equal(fused, again)}  nan={int(torch.isnan(fused).sum())}", flush=True)
    if shp is synth.TINY or B <= 3:
        p64 = orc.convdet_forward_f64(feat, w, b, shp.num_anchors, shp.num_fields).reshape(fused.shape)
        sc = np.abs(p64).mean()
        print(f"   vs float64: fused max {np.abs(fused.cpu().numpy() - p64).max() / sc:.2e}  staged max {np.abs(staged.cpu().numpy() - p64).max() / sc:.2e}", flush=True)
shp, B = synth.KITTI, 20
feats = [torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev)) for _ in range(3)]
w, b = synth.convdet_params(shp, 11); w, b = d(w), d(b)
packed = ops.pack_convdet_weights(w)
for name, val in (("fused", 1), ("staged", 0)):
    with _lib.option("SQD_HEAD_ONE_KERNEL", val):
        for i in range(5):
            ops.convdet_forward(feats[i % 3], w, b, packed=packed)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50):
            ops.convdet_forward(feats[i % 3], w, b, packed=packed)
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per convdet_forward (B=20)", flush=True)
