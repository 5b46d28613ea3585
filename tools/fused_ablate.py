"""Ablations of the one-kernel ConvDet path (SQD_F16_DBG bits: 1 no MMAs, 2 no feature loads, 4 no convert/stores, 8 no B loads)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import _lib, ops, synth
dev = torch.device("cuda")
shp, B = synth.KITTI, int(sys.argv[1]) if len(sys.argv) > 1 else 20
feats = [torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev)) for _ in range(3)]
w, b = synth.convdet_params(shp, 11)
w, b = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed = ops.pack_convdet_weights(w)
lib_ = _lib.load(); lib_.sqd_set_option(b"SQD_HEAD_ONE_KERNEL", 1)
for dbg in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,4,6,8,9,14,15".split(","))]:
    with _lib.option("SQD_F16_DBG", dbg):
        for i in range(5):
            ops.convdet_forward(feats[i % 3], w, b, packed=packed)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            ops.convdet_forward(feats[i % 3], w, b, packed=packed)
        e1.record(); torch.cuda.synchronize()
        print(f"dbg={dbg:2d}: {e0.elapsed_time(e1) / 30 * 1e3:.1f} us", flush=True)
if os.environ.get("SQD_LIB_PATH"):
    # profiling build: cycle accumulators per role (see PROF in convdet_fused.cu)
    names = {0: ("MMA", ["wait tempty", "wait afull", "wait bfull", "issue+commit"]),
             1: ("converter w0", ["wait afree", "-", "zero pads", "convert+store", "arrive+load issue", "loop top (geometry, data wait)"]),
             2: ("acc w0", ["wait tfull", "drain", "segment end / epilogue"]),
             3: ("B producer", ["wait bfree", "issue"])}
    for dbg in (0, 1):
        with _lib.option("SQD_F16_DBG", dbg):
            trace = torch.zeros((148, 4, 8), dtype=torch.int64, device=dev)
            os.environ["SQD_F16_TRACE"] = hex(trace.data_ptr())
            ops.convdet_forward(feats[0], w, b, packed=packed)
            torch.cuda.synchronize()
            del os.environ["SQD_F16_TRACE"]
            t = trace.cpu().numpy()
            print(f"---- dbg={dbg}: cycles per block (24.3 blocks per pair), CTA 0 / CTA 1 / mean over CTAs")
            for role, (rn, ks) in names.items():
                for k, kn in enumerate(ks):
                    col = t[:, role, k].astype(np.float64) / 24.3
                    nz = col[col > 0]
                    print(f"  {rn:13s} {kn:24s} {col[0]:8.0f} {col[1]:8.0f} {nz.mean() if nz.size else 0:8.0f}")
