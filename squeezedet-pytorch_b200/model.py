"""Drop-in mirrors of the reference's model classes (src/model/squeezedet.py) whose
post-backbone work runs in libsqdet_b200's CUDA kernels.

Same class names, constructor arguments (`cfg`), forward signatures, return types and state-dict
keys (`base.features.*`, `base.convdet.weight/bias`), so `utils/model.py:load_model` checkpoints
load unchanged and `eval.py` / `train.py` / `Detector` / `Trainer` call sites keep working.
The backbone (`Fire`, `features`) is OUT OF SCOPE of the path and stays stock PyTorch/cuDNN.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import CONV_SIMT_FP32, CONV_TCGEN05_F16X3, SqdError


class Fire(nn.Module):
    """SqueezeNet fire module (backbone, stock PyTorch).  squeezedet.py:9-23"""

    def __init__(self, inplanes, squeeze_planes, expand1x1_planes, expand3x3_planes):
        super().__init__()
        self.squeeze = nn.Conv2d(inplanes, squeeze_planes, kernel_size=1)
        self.expand1x1 = nn.Conv2d(squeeze_planes, expand1x1_planes, kernel_size=1)
        self.expand3x3 = nn.Conv2d(squeeze_planes, expand3x3_planes, kernel_size=3, padding=1)
        self.activation = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.activation(self.squeeze(x))
        return torch.cat([self.activation(self.expand1x1(x)), self.activation(self.expand3x3(x))], dim=1)


_ARCH = {
    # (stem conv, [fire / 'P' pool], convdet in-channels)      squeezedet.py:32-67
    "squeezedet": ((3, 64, 3, 2, 1), ["P", (64, 16, 64, 64), (128, 16, 64, 64), "P", (128, 32, 128, 128),
                                      (256, 32, 128, 128), "P", (256, 48, 192, 192), (384, 48, 192, 192),
                                      (384, 64, 256, 256), (512, 64, 256, 256), (512, 96, 384, 384),
                                      (768, 96, 384, 384)], 768),
    "squeezedetplus": ((3, 96, 7, 2, 3), ["P", (96, 96, 64, 64), (128, 96, 64, 64), (128, 192, 128, 128), "P",
                                          (256, 192, 128, 128), (256, 288, 192, 192), (384, 288, 192, 192),
                                          (384, 384, 256, 256), "P", (512, 384, 256, 256), (512, 384, 256, 256),
                                          (512, 384, 256, 256)], 512),
}


class _ConvDetFn(torch.autograd.Function):
    """ConvDet forward on the tcgen05 kernel.  Backward, all native: the feature gradient runs on the
    forward's tcgen05 kernel with swapped roles (ops.convdet_dgrad), the weight gradient on the tcgen05 pixel-contraction
    kernel (ops.convdet_wgrad; Cout <= 80 and an even grid width) or else on the fp32 CUDA-core implicit GEMM of the same
    library, the bias gradient on a cluster reduction.  There is no library (cuDNN / ATen) route: a shape the kernels
    do not take raises SqdError.

    grad_sink (dist.GradBucket or None): when a data-parallel bucket is attached, the weight / bias gradients are
    computed FIRST, added straight into the bucket's views and their all-reduce is launched before the feature
    gradient (and with it the whole backbone backward) is even enqueued, so the collective overlaps that work on the
    GPU ("kick NCCL from the wgrad epilogue").  Autograd then gets None for those two inputs."""

    @staticmethod
    def forward(ctx, x, weight, bias, packed, algo, dgrad_packed_fn, grad_sink=None):
        ctx.save_for_backward(x, weight, bias)
        ctx.dgrad_packed_fn = dgrad_packed_fn
        ctx.grad_sink = grad_sink
        return ops.convdet_forward(x, weight, bias, packed=packed, algo=algo)  # (B,gh,gw,Cout)

    @staticmethod
    def backward(ctx, g):
        x, weight, bias = ctx.saved_tensors
        gx = gw = gb = None
        g = g.contiguous()
        if ctx.needs_input_grad[0] and weight.shape[1] % 128 != 0:
            raise SqdError("ConvDet feature gradient: Cin %d is not a multiple of 128 (sqd_convdet_dgrad); detach the "
                           "features or use a backbone with 128-aligned Fire11 channels" % weight.shape[1])
        if ctx.needs_input_grad[1]:
            gw = ops.convdet_wgrad(x, g)
        if ctx.needs_input_grad[2]:
            gb = ops.convdet_bias_grad(g)
        sink = ctx.grad_sink
        if sink is not None and sink.take_early({id(weight): (weight, gw), id(bias): (bias, gb)}):
            gw = gb = None      # already in the bucket, their all-reduce is in flight
        if ctx.needs_input_grad[0]:
            gx = ops.convdet_dgrad(g, weight, ctx.dgrad_packed_fn() if ctx.dgrad_packed_fn else None)
        return gx, gw, gb, None, None, None, None


class SqueezeDetBase(nn.Module):
    """Backbone + ConvDet head -> pred (B, A, C+5).  squeezedet.py:26-97"""

    def __init__(self, cfg):
        super().__init__()
        self.num_classes = cfg.num_classes
        self.num_anchors = cfg.num_anchors
        if cfg.arch not in _ARCH:
            raise ValueError("Invalid architecture.")
        stem, body, head_in = _ARCH[cfg.arch]
        layers = [nn.Conv2d(stem[0], stem[1], kernel_size=stem[2], stride=stem[3], padding=stem[4]),
                  nn.ReLU(inplace=True)]
        for item in body:
            layers.append(nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True) if item == "P" else Fire(*item))
        self.features = nn.Sequential(*layers)
        self.dropout = nn.Dropout(cfg.dropout_prob, inplace=True) if cfg.dropout_prob > 0 else None
        self.convdet = nn.Conv2d(head_in, cfg.anchors_per_grid * (cfg.num_classes + 5), kernel_size=3, padding=1)
        self.conv_algo = getattr(cfg, "conv_algo", CONV_TCGEN05_F16X3)
        self._packed = None
        self._packed_version = None
        self._dgrad_packed = None
        self._dgrad_packed_version = None
        self.grad_sink = None     # dist.GradBucket of a data-parallel run (dist.bucket_for attaches it)
        self.init_weights()

    def init_weights(self):  # squeezedet.py:89-97
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, mean=0.0, std=0.002 if m is self.convdet else 0.005)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def packed_weights(self):
        """fp16 two-term split (w1, w2) weight planes for the tcgen05 f16x3 kernel; derived data, re-derived whenever the
        parameter changes (optimizer step, load_state_dict, .to(device)) and on every call while training (in-place
        updates through `.data` do not bump the version counter); never saved.  None for the SIMT algorithm."""
        if self.conv_algo == CONV_SIMT_FP32:
            return None
        w = self.convdet.weight
        ver = (w._version, w.data_ptr(), str(w.device))
        if self._packed is None or self._packed_version != ver or (self.training and torch.is_grad_enabled()):
            self._packed = ops.pack_convdet_weights(w)
            self._packed_version = ver
        return self._packed

    def invalidate_packed_weights(self):
        """Call after mutating convdet.weight through `.data` in eval mode (EMA, clipping ...): such writes bypass the
        version counter the caches are keyed on."""
        self._packed = self._dgrad_packed = None

    def dgrad_packed_weights(self):
        """Flipped / transposed planes for the feature gradient; derived lazily, only when a backward pass needs them."""
        w = self.convdet.weight
        ver = (w._version, w.data_ptr(), str(w.device))
        if self._dgrad_packed is None or self._dgrad_packed_version != ver or self.training:
            self._dgrad_packed = ops.pack_convdet_dgrad_weights(w)
            self._dgrad_packed_version = ver
        return self._dgrad_packed

    def head(self, feat):
        """ConvDet + permute/view of squeezedet.py:83-87 on a Fire11 feature map."""
        pred = _ConvDetFn.apply(feat, self.convdet.weight, self.convdet.bias, self.packed_weights(), self.conv_algo,
                                self.dgrad_packed_weights, self.grad_sink)
        return pred.view(-1, self.num_anchors, self.num_classes + 5)

    def forward(self, x):
        x = self.features(x)
        if self.dropout is not None:
            x = self.dropout(x)
        return self.head(x)


class PredictionResolver(nn.Module):
    """pred -> (class_probs, log_class_probs|None, scores(conf), deltas, boxes).  squeezedet.py:100-120.
    One fused kernel; the anchor table is a device buffer (the reference re-uploads it per call)."""

    def __init__(self, cfg, log_softmax=False):
        super().__init__()
        self.log_softmax = log_softmax
        self.input_size = cfg.input_size
        self.num_classes = cfg.num_classes
        self.anchors_per_grid = cfg.anchors_per_grid
        # (1, A, 4) like the reference's attribute (squeezedet.py:106); the kernels only use the memory
        self.register_buffer("anchors", torch.from_numpy(np.asarray(cfg.anchors)).float().contiguous().unsqueeze(0),
                             persistent=False)

    def _anchors_on(self, device):
        if self.anchors.device != device:
            self.anchors = self.anchors.to(device)
        return self.anchors

    def forward(self, pred):
        if pred.requires_grad and torch.is_grad_enabled():
            # the reference's resolver is differentiable; this one is the inference-side decode (one fused kernel, no
            # autograd graph).  Failing loudly beats silently returning zero gradients to a custom loss.
            raise SqdError("PredictionResolver.forward is inference-only here (its outputs carry no gradient): call it "
                           "under torch.no_grad() / on pred.detach(), or use Loss, which has its own native backward")
        want = ["probs", "conf", "deltas", "boxes"] + (["logp"] if self.log_softmax else [])
        out = ops.decode_scores(pred.detach(), self._anchors_on(pred.device), self.input_size, self.num_classes, want)
        return out["probs"], out.get("logp"), out["conf"], out["deltas"], out["boxes"]


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, anchors, input_size, num_classes, weights):
        losses, _ = ops.loss_fwd_bwd(pred, gt, anchors, input_size, num_classes, weights, want_grad=False)
        ctx.save_for_backward(pred, gt, anchors)
        ctx.meta = (input_size, num_classes, weights)
        return losses  # (B,4): class, positive score, negative score, bbox

    @staticmethod
    def backward(ctx, g):
        pred, gt, anchors = ctx.saved_tensors
        input_size, num_classes, weights = ctx.meta
        _, dpred = ops.loss_fwd_bwd(pred, gt, anchors, input_size, num_classes, weights, grad_loss=g.contiguous(),
                                    want_grad=True)
        return dpred, None, None, None, None, None


class Loss(nn.Module):
    """pred, gt -> (loss (B,), stats dict of (B,) vectors).  squeezedet.py:123-174"""

    def __init__(self, cfg):
        super().__init__()
        self.resolver = PredictionResolver(cfg, log_softmax=True)
        self.num_anchors = cfg.num_anchors
        self.num_classes = cfg.num_classes
        self.input_size = cfg.input_size
        self.class_loss_weight = cfg.class_loss_weight
        self.positive_score_loss_weight = cfg.positive_score_loss_weight
        self.negative_score_loss_weight = cfg.negative_score_loss_weight
        self.bbox_loss_weight = cfg.bbox_loss_weight

    def forward(self, pred, gt):
        weights = (self.class_loss_weight, self.positive_score_loss_weight, self.negative_score_loss_weight,
                   self.bbox_loss_weight)
        anchors = self.resolver._anchors_on(pred.device)
        t = _LossFn.apply(pred, gt, anchors, tuple(self.input_size), self.num_classes, weights)
        class_loss, pos, neg, bbox_loss = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        loss = class_loss + pos + neg + bbox_loss
        loss_stat = {"loss": loss, "class_loss": class_loss, "score_loss": pos + neg, "bbox_loss": bbox_loss}
        return loss, loss_stat


class SqueezeDetWithLoss(nn.Module):
    """Model for training.  squeezedet.py:177-187"""

    def __init__(self, cfg):
        super().__init__()
        self.base = SqueezeDetBase(cfg)
        self.loss = Loss(cfg)

    def forward(self, batch):
        pred = self.base(batch["image"])
        return self.loss(pred, batch["gt"])


class SqueezeDet(nn.Module):
    """Model for inference.  squeezedet.py:190-206"""

    def __init__(self, cfg):
        super().__init__()
        self.base = SqueezeDetBase(cfg)
        self.resolver = PredictionResolver(cfg, log_softmax=False)
        self.num_classes = cfg.num_classes
        self.input_size = cfg.input_size

    def forward(self, batch):
        pred = self.base(batch["image"])
        out = ops.decode_scores(pred.detach(), self.resolver._anchors_on(pred.device), self.input_size,
                                self.num_classes, ("class_ids", "scores", "boxes"))
        return {"class_ids": out["class_ids"], "scores": out["scores"], "boxes": out["boxes"]}
