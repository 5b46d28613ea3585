"""GPU check of the production tcgen05 ConvDet kernel (f16x3) against the SIMT fp32 kernel and a float64 torch
conv (all on the GPU) at several shapes and input magnitudes, plus its duration at the bench shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squeezedet_pytorch_b200 import ops, synth  # noqa: E402
from squeezedet_pytorch_b200._lib import CONV_SIMT_FP32, CONV_TCGEN05_F16X3, CONV_TCGEN05_F16X3_1CTA  # noqa: E402

TC = CONV_TCGEN05_F16X3_1CTA if "--1cta" in sys.argv else CONV_TCGEN05_F16X3


def run(shape, batch, layout, feat_scale=1.0, w_scale=1.0, ragged=False):
    g = torch.Generator(device="cuda").manual_seed(7)
    feat = torch.relu(torch.randn((batch, shape.in_channels, *shape.grid_hw), generator=g, device="cuda")) * feat_scale
    if ragged:  # very different magnitudes per image and per channel block
        feat = feat * torch.logspace(-6, 6, batch, device="cuda").view(batch, 1, 1, 1)
        feat[:, ::7] *= 1e-3
    if layout == "channels_last":
        feat = feat.contiguous(memory_format=torch.channels_last)
    w, b = synth.convdet_params(shape, 9)
    w, b = torch.from_numpy(w).cuda() * w_scale, torch.from_numpy(b).cuda()
    ref = torch.nn.functional.conv2d(feat.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).contiguous()
    simt = ops.convdet_forward(feat, w, b, algo=CONV_SIMT_FP32).double()
    tc = ops.convdet_forward(feat, w, b, algo=TC, check_status=True).double()
    den = ref.abs().flatten(1).mean(1).view(-1, 1, 1, 1)  # per-image typical magnitude
    e_tc, e_simt = ((tc - ref) / den).abs(), ((simt - ref) / den).abs()
    print(f"{shape.name:18s} B={batch:<4d} {layout:13s} fs={feat_scale:g} ws={w_scale:g} ragged={int(ragged)} | "
          f"f16x3 rel err max {float(e_tc.max()):.3e} rms {float(e_tc.pow(2).mean().sqrt()):.3e} "
          f"bias {float(((tc - ref) / den).mean()):+.2e} | simt max {float(e_simt.max()):.3e} "
          f"rms {float(e_simt.pow(2).mean().sqrt()):.3e}", flush=True)
    return feat, w, b


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    cases = ((synth.TINY, 1), (synth.TINY, 3), (synth.KITTI, 1), (synth.KITTI, 2), (synth.KITTI, 20), (synth.KITTI, 37),
             (synth.STRESS, 2))
    for shp, batch in cases[:3] if quick else cases:
        for layout in ("nchw", "channels_last"):
            run(shp, batch, layout)
    run(synth.KITTI, 3, "nchw", feat_scale=1e4, w_scale=1e-3)
    run(synth.KITTI, 3, "nchw", feat_scale=1e-5, w_scale=1e3)
    run(synth.KITTI, 5, "nchw", ragged=True)

    for layout in ("channels_last", "nchw"):
        feat, w, b = run(synth.KITTI, 20, layout)
        packed = ops.pack_convdet_weights(w)
        for _ in range(3):
            ops.convdet_forward(feat, w, b, packed=packed, algo=TC)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.convdet_forward(feat, w, b, packed=packed, algo=TC)
        e1.record()
        torch.cuda.synchronize()
        print(f"KITTI B=20 {layout:13s} f16x3: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per convdet_forward "
              f"(absmax + split + gemm)", flush=True)
