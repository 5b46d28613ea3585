"""GPU experiment: accuracy and duration of the tcgen05 ConvDet kernel as a function of the TMEM
accumulation chunk (SQD_TC_CHUNK pipeline stages of K=32 per epoch).  Error is measured against a
float64 evaluation on the CPU (2 KITTI images) -- the oracle is only the yardstick here."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from squeezedet_pytorch_b200 import ops, synth  # noqa: E402
from squeezedet_pytorch_b200._lib import CONV_SIMT_FP32, CONV_TCGEN05_F16X3  # noqa: E402

shp = synth.KITTI
feat = synth.features(shp, 2, 32)
w, b = synth.convdet_params(shp, 33)
p64 = orc.convdet_forward_f64(feat, w, b, shp.num_anchors, shp.num_fields)
ref32 = orc.convdet_forward(feat, w, b, shp.num_anchors, shp.num_fields)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
x, wd, bd = d(feat), d(w), d(b)
big = torch.relu(torch.randn((20, 768, 24, 78), device="cuda"))


def err(a):
    e = a.astype(np.float64) - p64
    return np.abs(e).max(), np.sqrt((e ** 2).mean()), float((e * np.sign(p64)).mean())


print("impl                      max|err|    rms        mean signed err (towards zero < 0)   ms @B=20")
print("torch CPU conv2d (fp32)   %.3e  %.3e  %+.3e" % err(ref32))
simt = ops.convdet_forward(x, wd, bd, algo=CONV_SIMT_FP32, num_fields=8).cpu().numpy()
print("SIMT fp32 FMA             %.3e  %.3e  %+.3e" % err(simt))
for chunk in (1, 2, 3, 4, 6, 8, 12, 24, 72, 216):
    os.environ["SQD_TC_CHUNK"] = str(chunk)
    out = ops.convdet_forward(x, wd, bd, algo=CONV_TCGEN05_F16X3, num_fields=8, check_status=True).cpu().numpy()
    packed = ops.pack_convdet_weights(wd)
    for _ in range(3):
        ops.convdet_forward(big, wd, bd, packed=packed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.convdet_forward(big, wd, bd, packed=packed)
    e1.record()
    torch.cuda.synchronize()
    print("tcgen05 3xTF32 chunk=%-4d %.3e  %.3e  %+.3e   %.3f (split+gemm)" % (chunk, *err(out), e0.elapsed_time(e1) / 10))
