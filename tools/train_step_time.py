"""Times the training-side mirror at BASELINE configs[3] (batch 20, KITTI): matcher + targets, ConvDet forward, loss
forward/backward, ConvDet backward (dgrad / wgrad / bias).  usage: python tools/train_step_time.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import ops, synth, targets  # noqa: E402

shp, B = synth.KITTI, 20
dev = torch.device("cuda")
feat = torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev))
w, b = synth.convdet_params(shp, 1)
w, b = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed, dpacked = ops.pack_convdet_weights(w), ops.pack_convdet_dgrad_weights(w)
anchors = synth.anchor_table(shp)
a32 = torch.from_numpy(anchors.astype(np.float32)).to(dev)
m = targets.AnchorMatcher(anchors, shp.num_classes)
cls_l, box_l = zip(*[synth.gt_boxes(shp, 100 + i) for i in range(B)])
gt_packed = m.pack(list(box_l), list(cls_l))


def timed(fn, n=20):
    for _ in range(3):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n * 1e3


gt, t_match = timed(lambda: m.dense_targets(*gt_packed))
pred, t_fwd = timed(lambda: ops.convdet_forward(feat, w, b, packed=packed, num_fields=shp.num_fields))
(losses, dpred), t_loss = timed(lambda: ops.loss_fwd_bwd(pred, gt, a32, shp.input_hw, shp.num_classes, (1.0, 3.75, 100.0, 6.0)))
g = dpred.view(B, *shp.grid_hw, shp.out_channels)
_, t_dgrad = timed(lambda: ops.convdet_dgrad(g, w, dpacked))
_, t_wgrad = timed(lambda: ops.convdet_wgrad(feat, g, tensor_cores=False))
_, t_wgrad_tc = timed(lambda: ops.convdet_wgrad(feat, g, tensor_cores=True))
_, t_bgrad = timed(lambda: ops.convdet_bias_grad(g))
gchw = g.permute(0, 3, 1, 2).contiguous()
torch.backends.cudnn.allow_tf32 = False
_, t_dgrad_t = timed(lambda: torch.nn.grad.conv2d_input(feat.shape, w, gchw, padding=1))
_, t_wgrad_t = timed(lambda: torch.nn.grad.conv2d_weight(feat, w.shape, gchw, padding=1))
print(f"batch {B} KITTI, us: matcher+targets {t_match:.1f} | convdet fwd {t_fwd:.1f} | loss fwd+bwd {t_loss:.1f} | "
      f"dgrad {t_dgrad:.1f} (cuDNN fp32 {t_dgrad_t:.1f}) | wgrad simt {t_wgrad:.1f} / tcgen05 {t_wgrad_tc:.1f} (cuDNN fp32 {t_wgrad_t:.1f}) | bias grad {t_bgrad:.1f}")
