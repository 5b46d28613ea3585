"""Deterministic synthetic inputs for the detection path.

All generators use ``numpy.random.RandomState`` (the frozen legacy stream), so a seed
reproduces the same bytes here, on the GPU box, and in ``oracle/gen_golden.py``; golden
fixtures therefore only need to store reference OUTPUTS.

The reference's own init (src/model/squeezedet.py:89-97: convdet std 0.002, zero bias) makes
every score tie at 1/(2C); the weights below are scaled so ConvDet logits have std ~ 1.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

F32 = np.float32

KITTI_SEEDS = np.array(
    [[34, 30], [75, 45], [38, 90], [127, 68], [80, 174], [196, 97], [194, 178], [283, 156], [381, 185]],
    dtype=np.float32,
)  # src/datasets/kitti.py:27-29


@dataclass(frozen=True)
class Shape:
    """Geometry of one workload: input size, grid (= size // 16, kitti.py:26), classes, filter knobs."""
    name: str
    input_hw: tuple
    num_classes: int
    top_k: int
    in_channels: int = 768
    anchors_per_grid: int = 9
    nms_thresh: float = 0.4
    score_thresh: float = 0.3

    @property
    def grid_hw(self):
        return (self.input_hw[0] // 16, self.input_hw[1] // 16)

    @property
    def num_anchors(self):
        return self.grid_hw[0] * self.grid_hw[1] * self.anchors_per_grid

    @property
    def num_fields(self):
        return self.num_classes + 5

    @property
    def out_channels(self):
        return self.anchors_per_grid * self.num_fields


KITTI = Shape("kitti_1248x384", (384, 1248), 3, 64)            # BASELINE.json configs[1..3]
STRESS = Shape("stress_2496x768", (768, 2496), 8, 256)          # BASELINE.json configs[4]
TINY = Shape("tiny_160x96", (96, 160), 3, 16)                   # oracle-speed parity case (6x10 grid)


def anchor_table(shape: Shape, seeds=KITTI_SEEDS) -> np.ndarray:
    """(A,4) float64 xywh anchor table, row a=(y*gw+x)*K+k (src/utils/boxes.py:37-67).

    Uses the reference's float64 expression for the centres so the table is bit-identical to
    the one the reference's matcher consumes."""
    gh, gw = shape.grid_hw
    ih, iw = shape.input_hw
    k = seeds.shape[0]
    cx = iw * (1 / (gw * 2) + np.linspace(0, 1, gw + 1)[:-1])
    cy = ih * (1 / (gh * 2) + np.linspace(0, 1, gh + 1)[:-1])
    t = np.empty((gh, gw, k, 4), dtype=np.float64)
    t[..., 0] = cx[None, :, None]
    t[..., 1] = cy[:, None, None]
    t[..., 2] = seeds[None, None, :, 0]
    t[..., 3] = seeds[None, None, :, 1]
    return t.reshape(-1, 4)


def features(shape: Shape, batch: int, seed: int) -> np.ndarray:
    """Post-ReLU Fire11-like feature map, (B,Cin,gh,gw) float32 NCHW."""
    gh, gw = shape.grid_hw
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((batch, shape.in_channels, gh, gw)).astype(F32)
    return np.maximum(x, F32(0))


def convdet_params(shape: Shape, seed: int):
    """ConvDet weight (K*(C+5),Cin,3,3) and bias, scaled for logit std ~ 1."""
    rs = np.random.RandomState(seed)
    fan = 9 * shape.in_channels
    w = (rs.standard_normal((shape.out_channels, shape.in_channels, 3, 3)) * (1.5 / np.sqrt(fan))).astype(F32)
    b = (rs.standard_normal((shape.out_channels,)) * 0.1).astype(F32)
    return w, b


def _sample_objects(rs, shape: Shape, n: int) -> np.ndarray:
    """KITTI-like boxes (n,4) float32 xyxy inside the image: w log-uniform 15..420 px (scaled to
    the image width), h = w * LogNormal(ln 0.55, 0.45)."""
    H, W = shape.input_hw
    sc = W / 1248.0
    w = np.exp(rs.uniform(np.log(15.0 * sc), np.log(420.0 * sc), size=n))
    h = w * np.exp(rs.normal(np.log(0.55), 0.45, size=n))
    w = np.minimum(w, W - 2.0)
    h = np.clip(h, 4.0, H - 2.0)
    cx = rs.uniform(w / 2, W - 1 - w / 2)
    cy = rs.uniform(h / 2, H - 1 - h / 2)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, W - 1)
    b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, H - 1)
    return b.astype(F32)


def gt_boxes(shape: Shape, seed: int, lo: int = 1, hi: int = 20):
    """Ground truth for one image: (class_ids (G,) int64, boxes (G,4) float32 xyxy), G~U{lo..hi},
    degenerate boxes dropped (the reference asserts x1<x2, y1<y2: src/utils/boxes.py:14-15)."""
    rs = np.random.RandomState(seed)
    g = int(rs.randint(lo, hi + 1))
    b = _sample_objects(rs, shape, g)
    ok = (b[:, 2] - b[:, 0] > 1) & (b[:, 3] - b[:, 1] > 1)
    b = b[ok]
    cls = rs.randint(0, shape.num_classes, size=b.shape[0]).astype(np.int64)
    return cls, b


def clustered_pred(shape: Shape, batch: int, seed: int, anchors=None) -> np.ndarray:
    """ConvDet-output-like tensor (B,A,C+5) float32 with real detection clusters.

    Background: class logits N(0,1), confidence logit N(-3,1), deltas N(0,0.3).  Per image
    3..12 planted objects; every anchor with IoU>0.3 to one gets confidence logit N(+3,1),
    +4 on the object's class logit and deltas that regress onto the object +N(0,0.05) -- so NMS
    has clusters to suppress (pure-random pred keeps 63/64)."""
    anchors = anchor_table(shape) if anchors is None else anchors
    A, C = shape.num_anchors, shape.num_classes
    rs = np.random.RandomState(seed)
    pred = np.empty((batch, A, C + 5), dtype=F32)
    pred[..., :C] = rs.standard_normal((batch, A, C))
    pred[..., C] = rs.standard_normal((batch, A)) - 3.0
    pred[..., C + 1:] = rs.standard_normal((batch, A, 4)) * 0.3
    ax1 = anchors[:, 0] - 0.5 * (anchors[:, 2] - 1)
    ay1 = anchors[:, 1] - 0.5 * (anchors[:, 3] - 1)
    ax2 = anchors[:, 0] + 0.5 * (anchors[:, 2] - 1)
    ay2 = anchors[:, 1] + 0.5 * (anchors[:, 3] - 1)
    aarea = (ax2 - ax1) * (ay2 - ay1)
    for b in range(batch):
        n_obj = int(rs.randint(3, 13))
        objs = _sample_objects(rs, shape, n_obj).astype(np.float64)
        ocls = rs.randint(0, C, size=n_obj)
        for o in range(n_obj):
            x1, y1, x2, y2 = objs[o]
            iw = np.maximum(np.minimum(ax2, x2) - np.maximum(ax1, x1), 0)
            ih = np.maximum(np.minimum(ay2, y2) - np.maximum(ay1, y1), 0)
            inter = iw * ih
            iou = inter / (aarea + (x2 - x1) * (y2 - y1) - inter + 1e-10)
            hit = np.nonzero(iou > 0.3)[0]
            if hit.size == 0:
                continue
            gx, gy, gw, gh = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1 + 1, y2 - y1 + 1
            pred[b, hit, C] = rs.standard_normal(hit.size) + 3.0
            pred[b, hit, ocls[o]] += 4.0
            reg = np.stack([
                (gx - anchors[hit, 0]) / anchors[hit, 2],
                (gy - anchors[hit, 1]) / anchors[hit, 3],
                np.log(gw / anchors[hit, 2]),
                np.log(gh / anchors[hit, 3]),
            ], axis=1)
            pred[b, hit, C + 1:] = reg + rs.standard_normal((hit.size, 4)) * 0.05
    return pred


def nonfinite_pred(shape: Shape, seed: int, anchors=None) -> np.ndarray:
    """clustered_pred for 5 images with NaN / +-inf planted in the CLASS and CONFIDENCE logits of the most confident
    anchors (where they change the result).  Deltas stay finite: the reference asserts on a non-finite decoded width
    (src/model/modules.py:18), so its behaviour is only defined for non-finite class / confidence logits."""
    pred = clustered_pred(shape, 5, seed, anchors=anchors)
    C = shape.num_classes
    top = [np.argsort(-pred[b, :, C], kind="stable")[:6] for b in range(5)]
    pred[0, top[0][0], :C + 1] = np.nan     # NaN class logits and confidence on the best anchor
    pred[0, top[0][3], C] = np.nan          # NaN confidence only
    pred[1, top[1][0], 1] = np.inf          # +inf class logit: softmax evaluates inf - inf = NaN
    pred[2, top[2][0], 0] = -np.inf         # -inf class logit: probability 0
    pred[2, top[2][1], :C] = -np.inf        # all -inf: NaN softmax
    pred[3, top[3][0], C] = np.inf          # confidence logit +inf: sigmoid 1
    pred[3, top[3][1], C] = -np.inf         # confidence logit -inf: sigmoid 0
    pred[4, top[4][2], C - 1] = np.nan      # a single NaN class logit
    return pred


def demo_state_dict(model, shape: Shape, seed: int):
    """Seeded weights for a whole SqueezeDet (backbone + ConvDet), keyed like `model.state_dict()` -- the reference's class
    and ours share the key names (utils/model.py:5-40).  The bundled checkpoint is absent and the reference's
    own init (std 0.005 / 0.002, squeezedet.py:89-97) collapses the Fire11 features to ~0, so the backbone convs get
    He-scaled normals (numpy RandomState, independent of torch's generator), zero biases, and the head gets
    convdet_params(): O(1) features, logits of std ~1, real detections.  Returns {name: torch tensor}."""
    import torch
    rs = np.random.RandomState(seed)
    out = {}
    for name, t in model.state_dict().items():
        shp = tuple(t.shape)
        if name.endswith("convdet.weight") or name.endswith("convdet.bias"):
            continue
        if len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            out[name] = torch.from_numpy((rs.standard_normal(shp) * np.sqrt(2.0 / fan_in)).astype(np.float32))
        else:
            out[name] = torch.zeros(shp, dtype=t.dtype)
    w, b = convdet_params(shape, seed + 1)
    for name in model.state_dict():
        if name.endswith("convdet.weight"):
            out[name] = torch.from_numpy(w)
        elif name.endswith("convdet.bias"):
            out[name] = torch.from_numpy(b)
    return out
