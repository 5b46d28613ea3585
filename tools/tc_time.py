"""Times the ConvDet GEMM alone (pre-split planes) at the bench shape under debug / tuning env settings.
usage: python tools/tc_time.py "SQD_F16_DBG=1" "SQD_F16_DBG=3,SQD_F16_A_STAGES=3" ..."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squeezedet_pytorch_b200 import _lib, ops, synth  # noqa: E402

ALGO = int(os.environ.get("TC_ALGO", "0"))
shp, B = synth.KITTI, int(os.environ.get("TC_BATCH", "20"))
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
feats = [torch.relu(torch.randn((B, shp.in_channels, *shp.grid_hw), generator=g, device=dev)) for _ in range(3)]
w, b = synth.convdet_params(shp, 9)
w, b = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed = ops.pack_convdet_weights(w)
gh, gw = shp.grid_hw
planes = [torch.empty(lib.sqd_convdet_split_bytes(B, shp.in_channels, gh, gw), dtype=torch.uint8, device=dev) for _ in range(3)]
st = _lib.stream_ptr(dev)
for f, pl in zip(feats, planes):
    _lib.check(lib.sqd_convdet_split_features(C.c_void_p(f.data_ptr()), 0, B, shp.in_channels, gh, gw, _lib.ptr(pl), st), "split")
ws = torch.empty(lib.sqd_convdet_workspace_bytes(B, shp.in_channels, gh, gw, shp.out_channels, 2, 0), dtype=torch.uint8, device=dev)
pred = torch.empty((B, gh, gw, shp.out_channels), device=dev)


def gemm(i):
    _lib.check(lib.sqd_convdet_forward(_lib.ptr(planes[i % 3]), 2, _lib.ptr(packed), None, _lib.ptr(b), B, shp.in_channels, gh, gw,
                                       shp.out_channels, _lib.ptr(pred), _lib.ptr(ws), ws.numel(), ALGO, st), "gemm")


for setting in (sys.argv[1:] or [""]):
    keys = []
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    for i in range(3):
        gemm(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        gemm(i)
    e1.record()
    torch.cuda.synchronize()
    rc = lib.sqd_convdet_status(_lib.ptr(ws), st)
    print(f"{setting or 'default':50s} {e0.elapsed_time(e1) / 30 * 1e3:8.1f} us per GEMM launch   status {rc}", flush=True)
    for k in keys:
        del os.environ[k]
