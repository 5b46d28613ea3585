"""Debug: per-unit pipeline timeline of one CTA of the tcgen05 ConvDet kernel (clock64 stamps).
usage: python tools/tc_trace.py [SQD_F16_DBG=7 ...]   (each argument = one env setting to trace under)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squeezedet_pytorch_b200 import _lib, ops, synth  # noqa: E402

ALGO = int(os.environ.get("TC_ALGO", "0"))
shp, B = synth.KITTI, 20
lib = _lib.load()
dev = torch.device("cuda")
feat = torch.relu(torch.randn((B, 768, 24, 78), device=dev))
w, b = synth.convdet_params(shp, 9)
w, b = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed = ops.pack_convdet_weights(w)
gh, gw = shp.grid_hw
planes = torch.empty(lib.sqd_convdet_split_bytes(B, 768, gh, gw), dtype=torch.uint8, device=dev)
st = _lib.stream_ptr(dev)
_lib.check(lib.sqd_convdet_split_features(C.c_void_p(feat.data_ptr()), 0, B, 768, gh, gw, _lib.ptr(planes), st), "split")
ws = torch.empty(lib.sqd_convdet_workspace_bytes(B, 768, gh, gw, 72, 2, 0), dtype=torch.uint8, device=dev)
pred = torch.empty((B, gh, gw, 72), device=dev)


def gemm():
    _lib.check(lib.sqd_convdet_forward(_lib.ptr(planes), 2, _lib.ptr(packed), None, _lib.ptr(b), B, 768, gh, gw, 72,
                                       _lib.ptr(pred), _lib.ptr(ws), ws.numel(), ALGO, st), "gemm")


NAMES = ["A_issue", "B_issue0", "m_tmem_ok", "m_a_ok", "m_b0_ok", "m_b0_iss", "m_b1_ok", "m_b1_iss", "m_b2_ok", "m_b2_iss",
         "m_end", "acc_start", "acc_done"]
for setting in (sys.argv[1:] or [""]):
    keys = []
    saved = {}
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        old = C.c_int(0)
        lib.sqd_get_option(k.encode(), C.byref(old))
        saved[k] = old.value
        lib.sqd_set_option(k.encode(), int(v))   # the option table is read from the environment only once
    for _ in range(3):
        gemm()
    trace = torch.zeros((512, 32), dtype=torch.int64, device=dev)
    os.environ["SQD_F16_TRACE"] = hex(trace.data_ptr())
    gemm()
    torch.cuda.synchronize()
    del os.environ["SQD_F16_TRACE"]
    t = trace.cpu().numpy()
    n = int((t[:, 10] > 0).sum())
    ce = t[:511, 12] > 0                      # units that closed a chunk (accumulate-warp stamps)
    if ce.sum() > 2:
        dr = (t[:511, 12] - t[:511, 11])[ce]
        print(f"==== {setting or 'default'}: accumulate warp 0 of the traced CTA: {int(ce.sum())} chunks, drain {dr[1:-1].mean():.0f} cycles (min {dr.min()}, max {dr.max()})")
    if n == 0:
        for k, v in saved.items():
            lib.sqd_set_option(k.encode(), v)
        continue
    t0 = t[0, 0]
    print(f"==== {setting or 'default'}: {n} units traced")
    ph = t[511, :8]
    if ph[0] > 0:
        names = ["entry", "setup done (barriers, bias, tensormap prefetch)", "tmem alloc + __syncthreads", "cluster sync", "pdl wait",
                 "producer warp done", "__syncthreads (all warps done)", "final cluster sync"]
        print("kernel phases of the traced CTA, cycles after entry: " + " | ".join(f"{nm} {int(v - ph[0])}" for nm, v in zip(names, ph)))
        print(f"first TMA issue {int(t0 - ph[0])} after entry; last acc_done {int(t[:n, 12].max() - ph[0])}")
    print("unit " + " ".join(f"{x:>9s}" for x in NAMES))
    for i in list(range(0, 5)) + list(range(36, 40)) + list(range(n - 3, n)):
        print(f"{i:4d} " + " ".join(f"{int(v - t0):9d}" for v in t[i][:13]))
    m = slice(2, n - 2)
    if ALGO == 0:
        d = lambda a, b: float((t[m, a] - t[m, b]).mean())  # noqa: E731
        print("pair: period %.0f | mma: wait tempty %.0f, wait full %.0f, issue+commit %.0f | producer issue -> full seen %.0f | "
              "acc: tfull seen after m_end %.0f, drain %.0f | tempty seen by mma(i+2) after acc_done(i) %.0f" % (
                  np.diff(t[:n, 10]).mean(), float((t[m, 2] - t[1:n - 3, 10]).mean()), d(3, 2), d(10, 3), d(3, 0), d(11, 10), d(12, 11),
                  float((t[4:n, 2] - t[2:n - 2, 12]).mean())))
        for k, v in saved.items():
            lib.sqd_set_option(k.encode(), v)
        continue
    print("period (m_end to m_end): mean %.0f median %.0f" % (np.diff(t[:n, 10]).mean(), np.median(np.diff(t[:n, 10]))))
    print("mma: wait tmem_empty %.0f | wait a_full %.0f | per dy: wait b_full %.0f %.0f %.0f, issue %.0f %.0f %.0f | tail commits %.0f | loop back %.0f" % (
        (t[m, 2] - t[1:n - 3, 10]).mean(), (t[m, 3] - t[m, 2]).mean(),
        (t[m, 4] - t[m, 3]).mean(), (t[m, 6] - t[m, 5]).mean(), (t[m, 8] - t[m, 7]).mean(),
        (t[m, 5] - t[m, 4]).mean(), (t[m, 7] - t[m, 6]).mean(), (t[m, 9] - t[m, 8]).mean(),
        (t[m, 10] - t[m, 9]).mean(), 0.0))
    print("acc: start after m_end %.0f | drain duration %.0f | idle between drains %.0f" % (
        (t[m, 11] - t[m, 10]).mean(), (t[m, 12] - t[m, 11]).mean(), (t[3:n - 1, 11] - t[m, 12]).mean()))
    print("A issue lead over m_a_ok %.0f | B issue(dy0) lead over m_b0_ok %.0f" % ((t[m, 3] - t[m, 0]).mean(), (t[m, 4] - t[m, 1]).mean()))
    for k, v in saved.items():
        lib.sqd_set_option(k.encode(), v)
