"""Build libsqdet_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed: the
library is a plain C ABI).  Usable as a module (`build()`) or a script."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libsqdet_b200_trace.so" if os.environ.get("SQD_BUILD_TRACE") else "libsqdet_b200.so")
SOURCES = ["api.cu", "io_kernels.cu", "decode.cu", "topk_nms.cu", "matcher.cu", "loss.cu", "convdet_simt.cu", "convdet_f16.cu", "convdet_fused.cu", "convdet_bwd.cu", "convdet_wgrad_tc.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# SQD_BUILD_TRACE=1: profiling build with the clock64 pipeline-trace stamps of the tcgen05 kernels compiled in
# (tools/tc_trace.py); the shipped library has none of it.
TRACE = ["-DSQD_ENABLE_TRACE"] if os.environ.get("SQD_BUILD_TRACE") else []
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build_trace" if TRACE else "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(HERE, "common.cuh"), os.path.join(HERE, "tc_ptx.cuh"), os.path.join(ROOT, "include", "sqdet_b200.h"), __file__]
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *FLAGS, *TRACE, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fvisibility=hidden"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
