"""CPU oracle for SqueezeDet's post-backbone detection path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package (``squeezedet-pytorch_b200``) never
imports anything from ``oracle/`` and fails loudly when its CUDA library is missing.

It is a restatement, in numpy, of the algorithm the reference runs for the path
SURVEY.md section 8 scopes (all citations relative to ``/root/reference``):

  generate_anchors        src/utils/boxes.py:37-67
  convdet_forward         src/model/squeezedet.py:73-75,83-87   (torch conv2d, see below)
  resolve / decode        src/model/squeezedet.py:109-120, src/model/modules.py:17-45,66-68
  score_argmax            src/model/squeezedet.py:200-205
  nms                     torchvision.ops.nms (third party, call site src/engine/detector.py:104)
  filter_image            src/engine/detector.py:87-122
  match_anchors           src/utils/boxes.py:70-135
  dense_targets           src/datasets/base.py:61-76
  pair_iou / loss         src/model/modules.py:48-63, src/model/squeezedet.py:133-174
  boxes_postprocess       src/utils/boxes.py:138-168

Third-party arithmetic that is NOT under /root/reference (reference pins in
requirements.txt:12,26,27; versions installed in this image in brackets):
  * torch==1.1.0 [2.11.0]: conv2d, exp, sigmoid, argsort, argmax, log_softmax.
    The conv is restated here as "call torch's CPU conv2d", which is exactly what the
    reference's nn.Conv2d does; everything else is restated in numpy.
  * torchvision==0.3.0 [0.26.0]: ops.nms.  Restated below from its published CPU
    algorithm (greedy, stable descending sort, strict '>' on float IoU promoted to
    double); pinned against the installed torchvision by oracle/gen_golden.py.
  * numpy==1.18.1 [2.3.5]: argsort in the matcher.  The reference's default
    (unstable) sort makes tie order implementation defined; the declared policy
    here is "lowest anchor index wins" == the reference with kind='stable'.

PARITY PINNING: the reference ships no tests or golden vectors.
The oracle is pinned by executing the reference's own modules in the build container
(oracle/gen_golden.py imports them from /root/reference/src) and committing their
outputs under tests/golden/; tests/test_oracle_golden.py replays them.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64

KITTI_SEEDS = np.array(
    [[34, 30], [75, 45], [38, 90], [127, 68], [80, 174], [196, 97], [194, 178], [283, 156], [381, 185]],
    dtype=np.float32,
)  # src/datasets/kitti.py:27-29


# --------------------------------------------------------------------------------------
# a10  anchor table
# --------------------------------------------------------------------------------------
def generate_anchors(grid_hw, input_hw, seeds):
    """(A,4) float64 xywh; row a = (y*gw + x)*K + k.  src/utils/boxes.py:37-67.

    Centres use the reference's exact float64 expression W*(1/(2*gw) + linspace) so that
    the table is bit-identical (the matcher consumes it in float64)."""
    gh, gw = int(grid_hw[0]), int(grid_hw[1])
    ih, iw = input_hw
    seeds = np.asarray(seeds)
    k = seeds.shape[0]
    cx = iw * (1 / (gw * 2) + np.linspace(0, 1, gw + 1)[:-1])
    cy = ih * (1 / (gh * 2) + np.linspace(0, 1, gh + 1)[:-1])
    out = np.empty((gh, gw, k, 4), dtype=F64)
    out[..., 0] = cx[None, :, None]
    out[..., 1] = cy[:, None, None]
    out[..., 2] = seeds[None, None, :, 0]
    out[..., 3] = seeds[None, None, :, 1]
    return out.reshape(-1, 4)


# --------------------------------------------------------------------------------------
# a1  ConvDet head
# --------------------------------------------------------------------------------------
def convdet_forward(feat_nchw, weight, bias, num_anchors, num_fields, threads=None):
    """feat (B,Cin,gh,gw) f32, weight (K*(C+5),Cin,3,3), bias -> pred (B,A,C+5) f32.

    src/model/squeezedet.py:83-87: conv2d(pad 1) -> permute(0,2,3,1) -> view(-1,A,C+5).
    The contraction itself is torch's CPU conv2d, as in the reference."""
    import torch

    if threads:
        torch.set_num_threads(int(threads))
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(feat_nchw, dtype=F32))
        w = torch.from_numpy(np.ascontiguousarray(weight, dtype=F32))
        b = torch.from_numpy(np.ascontiguousarray(bias, dtype=F32))
        y = torch.nn.functional.conv2d(x, w, b, stride=1, padding=1)
        y = y.permute(0, 2, 3, 1).contiguous().view(-1, num_anchors, num_fields)
    return y.numpy()


def convdet_forward_f64(feat_nchw, weight, bias, num_anchors, num_fields):
    """float64 direct evaluation of the same conv (small shapes only): the accuracy yardstick
    the float32 implementations (torch CPU, CUDA) are compared against."""
    x = np.asarray(feat_nchw, dtype=F64)
    w = np.asarray(weight, dtype=F64)
    B, Cin, H, W = x.shape
    Co = w.shape[0]
    xp = np.zeros((B, Cin, H + 2, W + 2), dtype=F64)
    xp[:, :, 1:-1, 1:-1] = x
    y = np.zeros((B, H, W, Co), dtype=F64)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + H, kx:kx + W]  # (B,Cin,H,W)
            y += np.einsum("bchw,oc->bhwo", patch, w[:, :, ky, kx], optimize=True)
    y += np.asarray(bias, dtype=F64)[None, None, None, :]
    return y.reshape(B, num_anchors, num_fields)


def convdet_backward(feat_nchw, weight, gpred_nhwc, dtype=F32):
    """Gradients of the ConvDet head for an upstream gradient of its (B,gh,gw,Cout) output: what autograd through
    nn.Conv2d(pad 1) + permute (src/model/squeezedet.py:83-85) returns -- (d feat (B,Cin,gh,gw), d weight, d bias).
    torch's CPU conv gradient routines do the contractions, as in the reference; dtype float64 = accuracy yardstick."""
    import torch

    td = torch.float64 if dtype == F64 else torch.float32
    x = torch.from_numpy(np.ascontiguousarray(feat_nchw)).to(td)
    w = torch.from_numpy(np.ascontiguousarray(weight)).to(td)
    g = torch.from_numpy(np.ascontiguousarray(gpred_nhwc)).to(td).permute(0, 3, 1, 2).contiguous()
    gx = torch.nn.grad.conv2d_input(x.shape, w, g, padding=1)
    gw = torch.nn.grad.conv2d_weight(x, w.shape, g, padding=1)
    gb = g.sum(dim=(0, 2, 3))
    return gx.numpy(), gw.numpy(), gb.numpy()


# --------------------------------------------------------------------------------------
# a2-a6  PredictionResolver
# --------------------------------------------------------------------------------------
def resolve(pred, anchors_xywh, input_hw, num_classes, log_softmax=False):
    """pred (B,A,C+5) f32 -> (probs, logp|None, conf, deltas, boxes), all f32.

    src/model/squeezedet.py:109-120; safe_softmax modules.py:66-68; deltas_to_boxes
    modules.py:27-45 (mul and add rounded separately, exp-scaled w/h, xywh->xyxy with the
    +-0.5*(w-1) convention of modules.py:17-24, clamp to [0,W-1] x [0,H-1])."""
    pred = np.asarray(pred, dtype=F32)
    C = num_classes
    z = pred[..., :C]
    zmax = z.max(axis=-1, keepdims=True)
    e = np.exp(z - zmax, dtype=F32)
    ssum = e[..., 0].copy()
    for c in range(1, C):  # left-to-right like a sequential reduce over the class axis
        ssum = ssum + e[..., c]
    probs = e / ssum[..., None]
    logp = None
    if log_softmax:
        logp = (z - zmax) - np.log(ssum, dtype=F32)[..., None]
    conf = (F32(1) / (F32(1) + np.exp(-pred[..., C:C + 1], dtype=F32))).astype(F32)
    deltas = np.ascontiguousarray(pred[..., C + 1:C + 5])
    anc = np.asarray(anchors_xywh).astype(F32)[None]  # .float() of the f64 table, squeezedet.py:106
    cx = anc[..., 0] + anc[..., 2] * deltas[..., 0]
    cy = anc[..., 1] + anc[..., 3] * deltas[..., 1]
    w = anc[..., 2] * np.exp(deltas[..., 2], dtype=F32)
    h = anc[..., 3] * np.exp(deltas[..., 3], dtype=F32)
    half = F32(0.5)
    one = F32(1)
    H, W = input_hw
    x1 = np.clip(cx - half * (w - one), F32(0), F32(W - 1))
    y1 = np.clip(cy - half * (h - one), F32(0), F32(H - 1))
    x2 = np.clip(cx + half * (w - one), F32(0), F32(W - 1))
    y2 = np.clip(cy + half * (h - one), F32(0), F32(H - 1))
    boxes = np.stack([x1, y1, x2, y2], axis=-1).astype(F32)
    return probs.astype(F32), logp, conf, deltas, boxes


def score_argmax(probs, conf):
    """src/model/squeezedet.py:200-205: probs *= conf; argmax (first max wins); max."""
    s = (probs * conf).astype(F32)
    ids = np.argmax(s, axis=-1).astype(np.int64)
    return ids, s.max(axis=-1)


def detect_dense(pred, anchors_xywh, input_hw, num_classes):
    """SqueezeDet.forward after the base: dict of class_ids i64 (B,A), scores (B,A), boxes (B,A,4)."""
    probs, _, conf, _, boxes = resolve(pred, anchors_xywh, input_hw, num_classes)
    ids, scores = score_argmax(probs, conf)
    return {"class_ids": ids, "scores": scores, "boxes": boxes}


# --------------------------------------------------------------------------------------
# a9  torchvision.ops.nms restated (CPU kernel semantics)
# --------------------------------------------------------------------------------------
def nms(boxes, scores, thresh):
    """Greedy NMS -> kept indices (into boxes) in descending-score order.

    torchvision CPU kernel rules: stable descending sort; area=(x2-x1)*(y2-y1) (no +1);
    inter = max(0,.)*max(0,.); iou = inter/((a_i+a_j)-inter) in float32; suppress when
    double(iou) > double(thresh) (strict); NaN never suppresses."""
    boxes = np.asarray(boxes, dtype=F32)
    scores = np.asarray(scores, dtype=F32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = desc_order(scores)
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = float(thresh)
    with np.errstate(invalid="ignore", divide="ignore"):
        for ii in range(n):
            i = order[ii]
            if dead[i]:
                continue
            keep.append(i)
            rest = order[ii + 1:]
            if rest.size == 0:
                continue
            w = np.maximum(F32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
            h = np.maximum(F32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
            inter = w * h
            iou = inter / ((areas[i] + areas[rest]) - inter)
            dead[rest[iou.astype(F64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


# --------------------------------------------------------------------------------------
# a8  Detector.filter with index tracking
# --------------------------------------------------------------------------------------
def desc_order(scores):
    """Stable descending order with torch's NaN rule: torch.sort / torch.argsort(descending=True) (detector.py:88 and
    the sort inside torchvision's nms kernel) treat NaN as LARGER than every number, +inf included, so NaN scores come
    first (numpy would put them last)."""
    scores = np.asarray(scores, dtype=F32)
    nan = np.isnan(scores)
    return np.lexsort((-np.where(nan, F32(0), scores), ~nan))   # primary: NaN first; secondary: score descending; stable


def topk_order(scores, k):
    """Declared tie policy: key = (score desc, anchor index asc); NaN scores rank first (torch)."""
    return desc_order(scores)[:k]


def filter_image(class_ids, scores, boxes, num_classes, top_k, nms_thresh, score_thresh):
    """One image of src/engine/detector.py:87-122, returning also the kept ANCHOR indices.

    -> dict(anchor_idx i64 (n,), class_ids i64, scores f32, boxes f32 (n,4)); n may be 0
    (the reference returns None in that case, detector.py:115-116).  Output order: classes
    ascending, each class in descending score, then the strict score > thresh filter."""
    class_ids = np.asarray(class_ids)
    scores = np.asarray(scores, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32)
    order = topk_order(scores, top_k)
    cid, sc, bx = class_ids[order], scores[order], boxes[order]
    out_idx = []
    for c in range(num_classes):
        sel = np.nonzero(cid == c)[0]
        if sel.size == 0:
            continue
        keep = nms(bx[sel], sc[sel], nms_thresh)
        out_idx.append(sel[keep])
    pos = np.concatenate(out_idx) if out_idx else np.zeros((0,), dtype=np.int64)
    pos = pos[sc[pos] > F32(score_thresh)]
    return {
        "anchor_idx": order[pos].astype(np.int64),
        "class_ids": cid[pos].astype(np.int64),
        "scores": sc[pos],
        "boxes": bx[pos],
    }


def detect_filtered(pred, anchors_xywh, input_hw, num_classes, top_k, nms_thresh, score_thresh):
    """pred (B,A,C+5) -> list of per-image filter_image dicts (the whole inference tail)."""
    det = detect_dense(pred, anchors_xywh, input_hw, num_classes)
    return [
        filter_image(det["class_ids"][b], det["scores"][b], det["boxes"][b], num_classes, top_k,
                     nms_thresh, score_thresh)
        for b in range(det["scores"].shape[0])
    ]


# --------------------------------------------------------------------------------------
# a11-a13  training-side matcher and dense targets
# --------------------------------------------------------------------------------------
def anchors_xyxy_f64(anchors_xywh):
    """numpy xywh_to_xyxy of src/utils/boxes.py:25-34 in float64."""
    a = np.asarray(anchors_xywh, dtype=F64)
    return np.stack([
        a[:, 0] - 0.5 * (a[:, 2] - 1),
        a[:, 1] - 0.5 * (a[:, 3] - 1),
        a[:, 0] + 0.5 * (a[:, 2] - 1),
        a[:, 1] + 0.5 * (a[:, 3] - 1),
    ], axis=1)


def match_anchors(gt_xyxy, anchors_xywh):
    """Greedy sequential anchor<->GT matching.  src/utils/boxes.py:84-135.

    gt_xyxy (G,4) float32 with x1<x2, y1<y2.  -> (deltas (G,4) f32, anchor_idx (G,) i32).
    For each GT box in annotation order: the untaken anchor with the largest IoU>0
    (float64 IoU, boxes.py:70-81; the GT area term is a float32 product promoted to double);
    if none, the untaken anchor nearest in squared xywh distance (boxes.py:115-121).
    Ties: lowest anchor index (== the reference with a stable argsort)."""
    gt = np.asarray(gt_xyxy, dtype=F32)
    anc = np.asarray(anchors_xywh, dtype=F64)
    A = anc.shape[0]
    axy = anchors_xyxy_f64(anc)
    area_a = (axy[:, 2] - axy[:, 0]) * (axy[:, 3] - axy[:, 1])
    taken = np.zeros(A, dtype=bool)
    G = gt.shape[0]
    idx = np.zeros((G,), dtype=np.int32)
    deltas = np.zeros((G, 4), dtype=F32)
    for i in range(G):
        b = gt[i]
        # float32 scalar arithmetic of the reference: (x1+x2)/2, x2-x1+1   (boxes.py:17-22)
        gx = (b[0] + b[2]) / F32(2)
        gy = (b[1] + b[3]) / F32(2)
        gw = b[2] - b[0] + F32(1)
        gh = b[3] - b[1] + F32(1)
        lr = np.maximum(np.minimum(axy[:, 2], F64(b[2])) - np.maximum(axy[:, 0], F64(b[0])), 0)
        tb = np.maximum(np.minimum(axy[:, 3], F64(b[3])) - np.maximum(axy[:, 1], F64(b[1])), 0)
        inter = lr * tb
        area_g = F64((b[2] - b[0]) * (b[3] - b[1]))  # float32 product, then promoted
        union = area_a + area_g - inter
        iou = inter / (union + 1e-10)
        cand = np.where(taken | ~(iou > 0), -np.inf, iou)
        j = int(np.argmax(cand))  # first (lowest index) maximum
        if not np.isfinite(cand[j]):
            d0 = F64(gx) - anc[:, 0]
            d1 = F64(gy) - anc[:, 1]
            d2 = F64(gw) - anc[:, 2]
            d3 = F64(gh) - anc[:, 3]
            dist = ((d0 * d0 + d1 * d1) + d2 * d2) + d3 * d3
            dist = np.where(taken, np.inf, dist)
            j = int(np.argmin(dist))
        taken[j] = True
        idx[i] = j
        deltas[i, 0] = (F64(gx) - anc[j, 0]) / anc[j, 2]
        deltas[i, 1] = (F64(gy) - anc[j, 1]) / anc[j, 3]
        deltas[i, 2] = np.log(F64(gw) / anc[j, 2])
        deltas[i, 3] = np.log(F64(gh) / anc[j, 3])
    return deltas, idx


def match_tie_audit(gt_xyxy, anchors_xywh, chosen):
    """Audit an anchor assignment `chosen` (G,) produced by ANY tie order of the reference's greedy matcher
    (src/utils/boxes.py:84-135; numpy's default argsort is unstable, so the unpatched reference breaks ties arbitrarily).

    Walks the boxes in annotation order with `chosen`'s own history of taken anchors and returns, per box,
    (valid, multiplicity): valid = the chosen anchor attains the best untaken IoU (or, with no overlapping untaken
    anchor, the smallest squared xywh distance); multiplicity = how many untaken anchors attain that optimum
    (1 = the choice was forced, no tie).  Same float64 arithmetic as match_anchors."""
    gt = np.asarray(gt_xyxy, dtype=F32)
    anc = np.asarray(anchors_xywh, dtype=F64)
    axy = anchors_xyxy_f64(anc)
    area_a = (axy[:, 2] - axy[:, 0]) * (axy[:, 3] - axy[:, 1])
    taken = np.zeros(anc.shape[0], dtype=bool)
    valid, mult = [], []
    for i in range(gt.shape[0]):
        b = gt[i]
        lr = np.maximum(np.minimum(axy[:, 2], F64(b[2])) - np.maximum(axy[:, 0], F64(b[0])), 0)
        tb = np.maximum(np.minimum(axy[:, 3], F64(b[3])) - np.maximum(axy[:, 1], F64(b[1])), 0)
        inter = lr * tb
        area_g = F64((b[2] - b[0]) * (b[3] - b[1]))
        iou = inter / (area_a + area_g - inter + 1e-10)
        cand = np.where(taken | ~(iou > 0), -np.inf, iou)
        j = int(chosen[i])
        best = cand.max()
        if np.isfinite(best):
            valid.append(bool(cand[j] == best))
            mult.append(int(np.count_nonzero(cand == best)))
        else:
            gx = (b[0] + b[2]) / F32(2)
            gy = (b[1] + b[3]) / F32(2)
            gw = b[2] - b[0] + F32(1)
            gh = b[3] - b[1] + F32(1)
            d0, d1, d2, d3 = F64(gx) - anc[:, 0], F64(gy) - anc[:, 1], F64(gw) - anc[:, 2], F64(gh) - anc[:, 3]
            dist = np.where(taken, np.inf, ((d0 * d0 + d1 * d1) + d2 * d2) + d3 * d3)
            best = dist.min()
            valid.append(bool(dist[j] == best))
            mult.append(int(np.count_nonzero(dist == best)))
        taken[j] = True
    return np.array(valid, dtype=bool), np.array(mult, dtype=np.int64)


def dense_targets(class_ids, gt_xyxy, anchors_xywh, num_classes):
    """src/datasets/base.py:61-76 -> gt (A, C+9) f32: [mask | box xyxy | deltas | one-hot]."""
    anc = np.asarray(anchors_xywh)
    deltas, idx = match_anchors(gt_xyxy, anc)
    gt = np.zeros((anc.shape[0], num_classes + 9), dtype=F32)
    gt[idx, 0] = 1.0
    gt[idx, 1:5] = np.asarray(gt_xyxy, dtype=F32)
    gt[idx, 5:9] = deltas
    gt[idx, 9 + np.asarray(class_ids, dtype=np.int64)] = 1.0
    return gt


# --------------------------------------------------------------------------------------
# a14-a16  loss forward and its analytic backward
# --------------------------------------------------------------------------------------
def pair_iou(b1, b2):
    """Elementwise IoU, src/model/modules.py:48-63 (float32, inter/(union+1e-10))."""
    b1 = np.asarray(b1, dtype=F32)
    b2 = np.asarray(b2, dtype=F32)
    lr = np.maximum(np.minimum(b1[..., 2], b2[..., 2]) - np.maximum(b1[..., 0], b2[..., 0]), F32(0))
    tb = np.maximum(np.minimum(b1[..., 3], b2[..., 3]) - np.maximum(b1[..., 1], b2[..., 1]), F32(0))
    inter = lr * tb
    union = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1]) + \
            (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1]) - inter
    return inter / (union + F32(1e-10))


def loss_forward(pred, gt, anchors_xywh, input_hw, num_classes, weights=(1.0, 3.75, 100.0, 6.0)):
    """src/model/squeezedet.py:133-174 -> dict of per-image (B,) f32 vectors
    {loss, class_loss, score_loss, bbox_loss} (+ positive/negative score parts)."""
    w_cls, w_pos, w_neg, w_box = (F32(w) for w in weights)
    pred = np.asarray(pred, dtype=F32)
    gt = np.asarray(gt, dtype=F32)
    A = pred.shape[1]
    m = gt[..., 0]
    gbox = gt[..., 1:5]
    gdel = gt[..., 5:9]
    onehot = gt[..., 9:]
    _, logp, conf, deltas, boxes = resolve(pred, anchors_xywh, input_hw, num_classes, log_softmax=True)
    conf = conf[..., 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        n = m.sum(axis=1, dtype=F32)
        iou = pair_iou(gbox, boxes) * m
        cls = (w_cls * m[..., None] * onehot * (-logp)).sum(axis=(1, 2), dtype=F32) / n
        pos = (w_pos * m * (iou - conf) ** 2).sum(axis=1, dtype=F32) / n
        neg = (w_neg * (F32(1) - m) * (iou - conf) ** 2).sum(axis=1, dtype=F32) / (F32(A) - n)
        box = (w_box * m[..., None] * (deltas - gdel) ** 2).sum(axis=(1, 2), dtype=F32) / n
    return {
        "loss": (cls + pos + neg + box).astype(F32),
        "class_loss": cls.astype(F32),
        "score_loss": (pos + neg).astype(F32),
        "bbox_loss": box.astype(F32),
        "positive_score_loss": pos.astype(F32),
        "negative_score_loss": neg.astype(F32),
    }


def loss_backward(pred, gt, anchors_xywh, input_hw, num_classes, grad_loss,
                  weights=(1.0, 3.75, 100.0, 6.0)):
    """Analytic d(sum_b grad_loss[b]*loss[b])/d pred, the gradient torch autograd produces for
    squeezedet.py:133-174 (IoU target NOT detached: gradient flows through the decoded box
    into the deltas; clamp passes gradient only where the raw coordinate is inside the
    image, modules.py:42-43; min/max ties split the gradient in half like torch)."""
    w_cls, w_pos, w_neg, w_box = (F64(w) for w in weights)
    pred = np.asarray(pred, dtype=F32)
    gt = np.asarray(gt, dtype=F32)
    B, A, _ = pred.shape
    C = num_classes
    H, W = input_hw
    m = gt[..., 0].astype(F64)
    g = gt[..., 1:5].astype(F64)
    gdel = gt[..., 5:9].astype(F64)
    onehot = gt[..., 9:].astype(F64)
    probs, _, conf, deltas, boxes = resolve(pred, anchors_xywh, input_hw, num_classes, log_softmax=True)
    probs = probs.astype(F64)
    sig = conf[..., 0].astype(F64)
    d = deltas.astype(F64)
    p = boxes.astype(F64)
    anc = np.asarray(anchors_xywh).astype(F32).astype(F64)[None]
    go = np.asarray(grad_loss, dtype=F64)
    if go.ndim == 1:                      # one upstream gradient per image, shared by the four terms
        go = np.repeat(go.reshape(B, 1), 4, axis=1)
    go_cls, go_pos, go_neg, go_box = (go[:, i:i + 1] for i in range(4))
    with np.errstate(invalid="ignore", divide="ignore"):
        n = m.sum(axis=1, keepdims=True)
        k_obj = m / n                         # per-anchor weight of the "/ num_objects" sums
        k_bg = (1.0 - m) / (A - n)
        # raw (unclamped) box, for the clamp pass-through mask
        cx = anc[..., 0] + anc[..., 2] * d[..., 0]
        cy = anc[..., 1] + anc[..., 3] * d[..., 1]
        bw = anc[..., 2] * np.exp(d[..., 2])
        bh = anc[..., 3] * np.exp(d[..., 3])
        raw = np.stack([cx - 0.5 * (bw - 1), cy - 0.5 * (bh - 1), cx + 0.5 * (bw - 1), cy + 0.5 * (bh - 1)], -1)
        lim = np.array([W - 1, H - 1, W - 1, H - 1], dtype=F64)
        passed = ((raw >= 0) & (raw <= lim)).astype(F64)
        # IoU and its partials w.r.t. the predicted box p
        lr_raw = np.minimum(g[..., 2], p[..., 2]) - np.maximum(g[..., 0], p[..., 0])
        tb_raw = np.minimum(g[..., 3], p[..., 3]) - np.maximum(g[..., 1], p[..., 1])
        lr = np.maximum(lr_raw, 0)
        tb = np.maximum(tb_raw, 0)
        inter = lr * tb
        pw = p[..., 2] - p[..., 0]
        ph = p[..., 3] - p[..., 1]
        union = (g[..., 2] - g[..., 0]) * (g[..., 3] - g[..., 1]) + pw * ph - inter
        den = union + 1e-10
        iou_raw = inter / den
        iou = iou_raw * m
        d_inter = (den + inter) / (den * den)
        d_area = -inter / (den * den)

        def sel(a, b):  # gradient share of `a` in min(a,b): 1, 0.5 on tie, 0
            return np.where(a < b, 1.0, np.where(a == b, 0.5, 0.0))

        on_lr = (lr_raw >= 0).astype(F64)
        on_tb = (tb_raw >= 0).astype(F64)
        dI = np.stack([
            -tb * on_lr * sel(-p[..., 0], -g[..., 0]),   # p.x1 enters through max(g.x1,p.x1)
            -lr * on_tb * sel(-p[..., 1], -g[..., 1]),
            tb * on_lr * sel(p[..., 2], g[..., 2]),
            lr * on_tb * sel(p[..., 3], g[..., 3]),
        ], -1)
        dA = np.stack([-ph, -pw, ph, pw], -1)
        diou_dp = d_inter[..., None] * dI + d_area[..., None] * dA
        resid = iou - sig
        k_sc = go_pos * w_pos * k_obj + go_neg * w_neg * k_bg
        dL_diou_raw = k_sc * 2.0 * resid * m
        Gp = dL_diou_raw[..., None] * diou_dp * passed
        grad = np.zeros((B, A, C + 5), dtype=F64)
        # class logits
        ysum = onehot.sum(-1, keepdims=True)
        grad[..., :C] = (go_cls * w_cls * k_obj)[..., None] * (ysum * probs - onehot)
        # confidence logit
        grad[..., C] = k_sc * 2.0 * resid * (-sig * (1.0 - sig))
        # deltas: bbox regression term + IoU path
        gd = (go_box * w_box * k_obj)[..., None] * 2.0 * (d - gdel)
        gd[..., 0] += anc[..., 2] * (Gp[..., 0] + Gp[..., 2])
        gd[..., 1] += anc[..., 3] * (Gp[..., 1] + Gp[..., 3])
        gd[..., 2] += 0.5 * bw * (Gp[..., 2] - Gp[..., 0])
        gd[..., 3] += 0.5 * bh * (Gp[..., 3] - Gp[..., 1])
        grad[..., C + 1:] = gd
    return grad.astype(F32)


# --------------------------------------------------------------------------------------
# 8(f) rank 1: boxes_postprocess
# --------------------------------------------------------------------------------------
def boxes_postprocess(boxes, image_meta):
    """src/utils/boxes.py:138-168: undo resize / pad / crop / flip / drift, in that order."""
    b = np.array(boxes, dtype=F32, copy=True)
    if "scales" in image_meta:
        b[:, [0, 2]] /= image_meta["scales"][1]
        b[:, [1, 3]] /= image_meta["scales"][0]
    if "padding" in image_meta:
        b[:, [0, 2]] -= image_meta["padding"][2]
        b[:, [1, 3]] -= image_meta["padding"][0]
    if "crops" in image_meta:
        b[:, [0, 2]] += image_meta["crops"][2]
        b[:, [1, 3]] += image_meta["crops"][0]
    if image_meta.get("flipped", False):
        width = image_meta["drifted_size"][1] if "drifted_size" in image_meta else image_meta["orig_size"][1]
        bw = b[:, 2] - b[:, 0] + F32(1)
        b[:, 0] = width - 1 - b[:, 2]
        b[:, 2] = b[:, 0] + bw - F32(1)
    if "drifts" in image_meta:
        b[:, [0, 2]] += image_meta["drifts"][1]
        b[:, [1, 3]] += image_meta["drifts"][0]
    return b


# --------------------------------------------------------------------------------------
# 8(f) rank 3: KITTI result lines
# --------------------------------------------------------------------------------------
def kitti_result_text(class_ids, scores, boxes, class_names):
    """src/datasets/kitti.py:88-96: the content of one image's result file."""
    out = []
    for i in range(len(class_ids)):
        out.append("{} -1 -1 0 {:.2f} {:.2f} {:.2f} {:.2f} 0 0 0 0 0 0 0 {:.3f}\n".format(
            class_names[int(class_ids[i])].lower(), *boxes[i, :], scores[i]))
    return "".join(out)


# --------------------------------------------------------------------------------------
# 8(f) rank 4: whiten + resize (+ HWC -> CHW)
# --------------------------------------------------------------------------------------
def preprocess_image(image_hwc, mean, std, out_hw):
    """src/utils/image.py:9-19 (whiten) then :77-88 (cv2.resize, default INTER_LINEAR) then base.py:33 (transpose).
    cv2's float32 bilinear restated: fx = float((d + .5) * scale - .5), s = floor, weights (1 - f, f), borders clamp
    with a zero weight; horizontal pass per source row, then the vertical pass."""
    img = (np.asarray(image_hwc, dtype=F32) - np.asarray(mean, F32).reshape(1, 1, 3)) / np.asarray(std, F32).reshape(1, 1, 3)
    H0, W0 = img.shape[:2]
    H, W = out_hw

    def table(n_dst, n_src):
        scale = 1.0 / (float(n_dst) / n_src)
        d = np.arange(n_dst, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(F32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(F32)).astype(F32)
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= n_src - 1
        f[hi], s[hi] = 0, n_src - 1
        return s, np.minimum(s + 1, n_src - 1), f

    x0, x1, fx = table(W, W0)
    y0, y1, fy = table(H, H0)
    a0 = (F32(1) - fx)[None, :, None]
    rows = img[:, x0, :] * a0 + img[:, x1, :] * fx[None, :, None]           # (H0, W, 3)
    out = rows[y0] * (F32(1) - fy)[:, None, None] + rows[y1] * fy[:, None, None]
    return np.ascontiguousarray(out.astype(F32).transpose(2, 0, 1))
