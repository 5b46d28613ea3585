"""Can the feature pre-pass of the NEXT step run on the SMs while the ConvDet GEMM of the current step holds them?
GEMM-only launches (pre-split planes) on stream A, pre-pass launches on stream B: alone, and together.
usage: python tools/coresidency_probe.py [reps]     (SQD_SPLIT_TWO_PASS=1 selects the light two-kernel pre-pass)"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import _lib, ops, synth
lib = _lib.load()
dev = torch.device("cuda")
shp, B = synth.KITTI, 20
reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 50
gh, gw = shp.grid_hw
cin, cout = shp.in_channels, shp.out_channels
feats = [torch.relu(torch.randn((B, cin, gh, gw), device=dev)) for _ in range(2)]
w, b = synth.convdet_params(shp, 4321)
weight, bias = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed = ops.pack_convdet_weights(weight)
pbytes = lib.sqd_convdet_split_bytes(B, cin, gh, gw)
planes = [torch.empty(pbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
pA, pB = C.c_void_p(sA.cuda_stream), C.c_void_p(sB.cuda_stream)
_lib.check(lib.sqd_convdet_split_features(C.c_void_p(feats[0].data_ptr()), _lib.LAYOUT_NCHW, B, cin, gh, gw, _lib.ptr(planes[0]), pA), "split")
ws = torch.empty(lib.sqd_convdet_workspace_bytes(B, cin, gh, gw, cout, _lib.LAYOUT_SPLIT_NHWC, _lib.CONV_TCGEN05_F16X3), dtype=torch.uint8, device=dev)
pred = torch.empty((B, gh * gw * shp.anchors_per_grid, shp.num_classes + 5), device=dev)


def gemm():
    _lib.check(lib.sqd_convdet_forward(_lib.ptr(planes[0]), _lib.LAYOUT_SPLIT_NHWC, _lib.ptr(packed), None, _lib.ptr(bias), B, cin, gh, gw,
                                       cout, _lib.ptr(pred), _lib.ptr(ws), ws.numel(), _lib.CONV_TCGEN05_F16X3, pA), "gemm")


big = torch.empty((B * cin * gh * gw,), device=dev)


def prepass():
    if "--mul" in sys.argv:          # any light kernel: an ATen elementwise multiply over 115 MB
        with torch.cuda.stream(sB):
            big.mul_(1.0001)
        return
    _lib.check(lib.sqd_convdet_split_features(C.c_void_p(feats[1].data_ptr()), _lib.LAYOUT_NCHW, B, cin, gh, gw, _lib.ptr(planes[1]), pB), "split")


def run(do_a, do_b):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    sA.wait_event(e0); sB.wait_event(e0)
    for _ in range(reps):
        if do_a:
            gemm()
        if do_b:
            prepass()
    main.wait_stream(sA); main.wait_stream(sB)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for _ in range(2):
    run(True, True)
a, b_, ab = run(True, False), run(False, True), run(True, True)
print(f"per pair of launches: GEMM alone {a:.1f} us | pre-pass alone {b_:.1f} us | both streams together {ab:.1f} us "
      f"(sum {a + b_:.1f}, max {max(a, b_):.1f}) -> overlap {(a + b_ - ab) / min(a, b_) * 100:.0f} % of the shorter one")
