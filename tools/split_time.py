"""Times the feature pre-pass (sqd_convdet_split_features) alone for a few kernel configurations (selected through
sqd_set_option: the option table is read from the environment only once).  usage: python tools/split_time.py [batch]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import _lib, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = _lib.load()
shp = synth.KITTI
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)
feats = [torch.relu(torch.randn((B, shp.in_channels, *shp.grid_hw), generator=gen, device=dev)) for _ in range(3)]
planes = torch.empty(lib.sqd_convdet_split_bytes(B, shp.in_channels, *shp.grid_hw), dtype=torch.uint8, device=dev)
st = _lib.stream_ptr(dev)
nbytes = feats[0].numel() * 8


DEFAULTS = {"SQD_SPLIT_TWO_PASS": 0, "SQD_SPLIT_CS": 0, "SQD_SPLIT_THREADS": 512, "SQD_SPLIT_ROWS": 0, "SQD_SPLIT_REGS": 0}


def run(tag, env):
    for k, v in {**DEFAULTS, **{k: int(v) for k, v in env.items()}}.items():
        _lib.check(lib.sqd_set_option(k.encode(), int(v)), "sqd_set_option")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    for i in range(n + 5):
        if i == 5:
            e0.record()
        _lib.check(lib.sqd_convdet_split_features(C.c_void_p(feats[i % 3].data_ptr()), 0, B, shp.in_channels, shp.grid_hw[0],
                                                  shp.grid_hw[1], _lib.ptr(planes), st), "split")
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{tag:40s} {us:8.1f} us  {nbytes / us / 1e6:6.2f} TB/s (read+write)")


run("default (one-pass cluster kernel, shared tile)", {})
run("register-resident one-pass (SQD_SPLIT_REGS)", {"SQD_SPLIT_REGS": "1"})
run("two-pass (absmax + split)", {"SQD_SPLIT_TWO_PASS": "1"})
for cs in (8, 16):
    for th in (256, 512, 1024):
        for rows in (0, 1):
            run(f"one-pass cs={cs} threads={th} rows={rows}", {"SQD_SPLIT_CS": str(cs), "SQD_SPLIT_THREADS": str(th), "SQD_SPLIT_ROWS": str(rows)})
