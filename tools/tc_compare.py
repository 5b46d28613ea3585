"""GPU check of the production tcgen05 ConvDet kernel against the v1 kernel and the SIMT fp32 kernel
(all on the GPU) at several shapes, plus its duration at the bench shape."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squeezedet_pytorch_b200 import ops, synth  # noqa: E402
from squeezedet_pytorch_b200._lib import CONV_SIMT_FP32, CONV_TCGEN05_3XTF32, CONV_TCGEN05_V1, CONV_TCGEN05_V2  # noqa: E402


def run(shape, batch, layout):
    g = torch.Generator(device="cuda").manual_seed(7)
    feat = torch.relu(torch.randn((batch, shape.in_channels, *shape.grid_hw), generator=g, device="cuda"))
    if layout == "channels_last":
        feat = feat.contiguous(memory_format=torch.channels_last)
    w, b = synth.convdet_params(shape, 9)
    w, b = torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda()
    simt = ops.convdet_forward(feat, w, b, algo=CONV_SIMT_FP32)
    v2 = ops.convdet_forward(feat, w, b, algo=CONV_TCGEN05_3XTF32, check_status=True)
    d2 = (v2 - simt).abs()
    msg = f"{shape.name:18s} B={batch:<4d} {layout:13s} v3-simt max {float(d2.max()):.3e} mean {float(d2.mean()):.3e}"
    if shape.out_channels <= 128:
        v1 = ops.convdet_forward(feat, w, b, algo=CONV_TCGEN05_V1, check_status=True)
        msg += f" | v1-simt max {float((v1 - simt).abs().max()):.3e}"
    print(msg, flush=True)
    return feat, w, b


for shp, batch in ((synth.TINY, 1), (synth.TINY, 3), (synth.KITTI, 1), (synth.KITTI, 2), (synth.KITTI, 20),
                   (synth.KITTI, 37), (synth.STRESS, 2)):
    for layout in ("nchw", "channels_last"):
        run(shp, batch, layout)

for layout in ("channels_last", "nchw"):
    feat, w, b = run(synth.KITTI, 20, layout)
    packed = ops.pack_convdet_weights(w)
    for algo, name in ((CONV_TCGEN05_3XTF32, "v3"), (CONV_TCGEN05_V2, "v2"), (CONV_TCGEN05_V1, "v1")):
        for _ in range(3):
            ops.convdet_forward(feat, w, b, packed=packed, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.convdet_forward(feat, w, b, packed=packed, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        print(f"KITTI B=20 {layout:13s} {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per convdet_forward", flush=True)
