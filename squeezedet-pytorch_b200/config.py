"""The `cfg` namespace the reference's classes read (src/utils/config.py:121-131).  Our modules
take the same object with the same field names, so a cfg built by the reference's Config /
update_dataset_info works unchanged; these helpers build one without the dataset classes."""
from __future__ import annotations

import types

import numpy as np

from . import synth


def make_config(shape: synth.Shape = synth.KITTI, device="cuda", **overrides):
    """Namespace with every hot-path field: squeezedet.py:29-30,71-75,104-107,127-131 and
    detector.py:16,42-48,88,95,104,114.  Defaults are the reference's (config.py:10-85,
    kitti.py:15-32)."""
    anchors = synth.anchor_table(shape)
    cfg = types.SimpleNamespace(
        arch="squeezedet", dropout_prob=0.5,
        input_size=tuple(shape.input_hw), num_classes=shape.num_classes,
        class_names=("Car", "Pedestrian", "Cyclist") if shape.num_classes == 3 else
        tuple(f"class_{i}" for i in range(shape.num_classes)),
        anchors=anchors, anchors_per_grid=shape.anchors_per_grid, num_anchors=anchors.shape[0],
        grid_size=tuple(shape.grid_hw), anchors_seed=synth.KITTI_SEEDS,
        rgb_mean=np.array([93.877, 98.801, 95.923], dtype=np.float32).reshape(1, 1, 3),
        rgb_std=np.array([78.782, 80.130, 81.200], dtype=np.float32).reshape(1, 1, 3),
        keep_top_k=shape.top_k, nms_thresh=shape.nms_thresh, score_thresh=shape.score_thresh,
        class_loss_weight=1.0, positive_score_loss_weight=3.75, negative_score_loss_weight=100.0,
        bbox_loss_weight=6.0,
        device=device, debug=0, mode="eval", batch_size=20, num_workers=0, print_interval=10,
        debug_dir="debug", gpus=[0],
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def kitti_config(device="cuda", **overrides):
    return make_config(synth.KITTI, device=device, **overrides)
