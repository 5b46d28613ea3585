"""Training-side targets: generate_anchors / compute_deltas / prepare_annotations with the
reference's signatures (src/utils/boxes.py:37-135, src/datasets/base.py:61-76), executed by the
float64 CUDA matcher.  Batched variants keep everything on the device."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def generate_anchors(grid_size, input_size, anchors_seed):
    """(A,4) float64 xywh table, row a=(y*gw+x)*K+k -- host-side, once at start-up, bit-identical
    to src/utils/boxes.py:37-67 (same float64 expression for the centres)."""
    gh, gw = int(grid_size[0]), int(grid_size[1])
    ih, iw = input_size
    seeds = np.asarray(anchors_seed)
    assert seeds.ndim == 2 and seeds.shape[1] == 2
    cx = iw * (1 / (gw * 2) + np.linspace(0, 1, gw + 1)[:-1])
    cy = ih * (1 / (gh * 2) + np.linspace(0, 1, gh + 1)[:-1])
    t = np.empty((gh, gw, seeds.shape[0], 4), dtype=np.float64)
    t[..., 0] = cx[None, :, None]
    t[..., 1] = cy[:, None, None]
    t[..., 2] = seeds[None, None, :, 0]
    t[..., 3] = seeds[None, None, :, 1]
    return t.reshape(-1, 4)


class AnchorMatcher:
    """Device-resident anchor table + batched matcher / dense-target builder."""

    def __init__(self, anchors_xywh, num_classes, device="cuda"):
        a = np.ascontiguousarray(np.asarray(anchors_xywh, dtype=np.float64))
        assert a.ndim == 2 and a.shape[1] == 4
        self.num_anchors = a.shape[0]
        self.num_classes = int(num_classes)
        self.device = torch.device(device)
        self.anchors64 = torch.from_numpy(a).to(self.device)

    @staticmethod
    def _check(boxes):
        # the reference asserts x1<x2 and y1<y2 (src/utils/boxes.py:14-15)
        assert boxes.ndim == 2 and boxes.shape[1] == 4
        assert np.all(boxes[:, 0] < boxes[:, 2]) and np.all(boxes[:, 1] < boxes[:, 3])

    def pack(self, boxes_list, classes_list=None):
        """list of (G_i,4) float32 arrays -> padded device tensors (boxes, classes, count)."""
        B = len(boxes_list)
        gmax = max(1, max((len(b) for b in boxes_list), default=1))
        boxes = np.zeros((B, gmax, 4), dtype=np.float32)
        classes = np.zeros((B, gmax), dtype=np.int32)
        count = np.zeros((B,), dtype=np.int32)
        for i, b in enumerate(boxes_list):
            b = np.asarray(b, dtype=np.float32).reshape(-1, 4)
            self._check(b)
            if b.shape[0] > self.num_anchors:
                raise IndexError("more ground-truth boxes than anchors")  # the reference fails here too
            boxes[i, :len(b)] = b
            count[i] = len(b)
            if classes_list is not None:
                classes[i, :len(b)] = np.asarray(classes_list[i], dtype=np.int32)
        dev = self.device
        return torch.from_numpy(boxes).to(dev), torch.from_numpy(classes).to(dev), torch.from_numpy(count).to(dev)

    def match(self, gt_boxes, gt_count):
        return ops.match_anchors(gt_boxes, gt_count, self.anchors64)

    def dense_targets(self, gt_boxes, gt_classes, gt_count):
        """-> gt (B, A, C+9) on the device, the tensor SqueezeDetWithLoss consumes as batch['gt']."""
        idx, deltas = self.match(gt_boxes, gt_count)
        return ops.build_targets(gt_boxes, gt_classes, gt_count, idx, deltas, self.num_anchors, self.num_classes)


_matchers = {}


def _matcher_for(anchors_xywh, num_classes, device):
    a = np.asarray(anchors_xywh)
    key = (a.shape, a[:4].tobytes(), a[-4:].tobytes(), int(num_classes), str(device))
    m = _matchers.get(key)
    if m is None:
        m = _matchers[key] = AnchorMatcher(a, num_classes, device)
    return m


def compute_deltas(boxes_xyxy, anchors_xywh, device="cuda"):
    """Drop-in for src/utils/boxes.py:84-135: (G,4) float32 xyxy, (A,4) xywh ->
    (deltas (G,4) float32, anchor_indices (G,) int32), numpy in / numpy out."""
    boxes = np.asarray(boxes_xyxy, dtype=np.float32)
    m = _matcher_for(anchors_xywh, 1, device)
    if boxes.shape[0] == 0:
        return np.zeros((0, 4), np.float32), np.zeros((0,), np.int32)
    gb, _, gc = m.pack([boxes])
    idx, deltas = m.match(gb, gc)
    g = boxes.shape[0]
    return deltas[0, :g].cpu().numpy(), idx[0, :g].cpu().numpy()


def prepare_annotations(class_ids, boxes, anchors_xywh, num_classes, device="cuda"):
    """Drop-in for BaseDataset.prepare_annotations (src/datasets/base.py:61-76) -> (A, C+9) float32."""
    boxes = np.asarray(boxes, dtype=np.float32)
    m = _matcher_for(anchors_xywh, num_classes, device)
    gb, gcl, gc = m.pack([boxes], [class_ids])
    return m.dense_targets(gb, gcl, gc)[0].cpu().numpy()
