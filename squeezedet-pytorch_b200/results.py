"""Result side of the path: one packed device->host transfer per batch and the KITTI
evaluator's text files.  Mirrors KITTI.save_results (src/datasets/kitti.py:78-97) and the result half of
Detector.detect (src/engine/detector.py:33-40)."""
from __future__ import annotations

import os

import torch

from . import ops

KITTI_CLASS_NAMES = ("Car", "Pedestrian", "Cyclist")   # src/datasets/kitti.py:16


def to_host(det: ops.Detections, meta: torch.Tensor = None):
    """Detections (device) -> (packed (B,k,6) float32, count (B,) int32) on the host; boxes post-processed on the
    device when `meta` ((B,10) records, see sqd_boxes_postprocess) is given.  Two copies for the whole batch."""
    packed = ops.pack_results(det, meta)
    return packed.cpu(), det.count.cpu()


def unpack(packed_host, count_host, image_metas=None):
    """-> the reference's list of per-image dicts (numpy views into the packed array), detector.py:33-40."""
    p, counts = packed_host.numpy(), count_host.tolist()
    out = []
    for b, n in enumerate(counts):
        meta = image_metas[b] if image_metas is not None else {}
        if n == 0:
            out.append({"image_meta": meta})
            continue
        out.append({"class_ids": p[b, :n, 0].astype("int64"), "scores": p[b, :n, 1], "boxes": p[b, :n, 2:6],
                    "image_meta": meta})
    return out


def kitti_texts(packed_host, count_host, class_names=KITTI_CLASS_NAMES):
    """One string per image: the content of its KITTI result file (empty when nothing was kept)."""
    return ops.format_kitti(packed_host, count_host, class_names)


def save_results(results_dir, image_ids, packed_host, count_host, class_names=KITTI_CLASS_NAMES):
    """kitti.py:78-97: <results_dir>/data/<image_id>.txt for every image of the batch."""
    txt_dir = os.path.join(results_dir, "data")
    os.makedirs(txt_dir, exist_ok=True)
    for image_id, text in zip(image_ids, kitti_texts(packed_host, count_host, class_names)):
        with open(os.path.join(txt_dir, str(image_id) + ".txt"), "w") as fp:
            fp.write(text)
