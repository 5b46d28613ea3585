"""GPU parity tests proper: every CUDA entry point (through the C ABI, via the ctypes host layer)
against the CPU oracle on the same seeded inputs, against the golden fixtures recorded from the
reference, and -- at BASELINE.json's full sizes -- through size-independent properties.

Bars: indices / class ids / counts bit-exact; scores, boxes, losses, gradients within 1e-4
relative (north_star), with the absolute floors of SURVEY 8c."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from squeezedet_pytorch_b200 import _lib, synth
from conftest import split_ragged

pytestmark = pytest.mark.gpu

RTOL = 1e-4
SHAPES = {s.name: s for s in (synth.TINY, synth.KITTI, synth.STRESS)}


@pytest.fixture(scope="module")
def ops():
    from squeezedet_pytorch_b200 import ops as _ops
    return _ops


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def anchors_dev(shape):
    a = synth.anchor_table(shape)
    return a, dev(a.astype(np.float32))


def dets_to_lists(det):
    return det.to_list()


# ------------------------------------------------------------------------------------------------------
# a2-a7 decode
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,batch", [("tiny_160x96", 4), ("kitti_1248x384", 2), ("stress_2496x768", 1)])
def test_decode_vs_oracle(ops, name, batch):
    shp = SHAPES[name]
    a64, a32 = anchors_dev(shp)
    pred = synth.clustered_pred(shp, batch, 101, anchors=a64)
    out = ops.decode_scores(dev(pred), a32, shp.input_hw, shp.num_classes,
                            ("class_ids", "scores", "boxes", "probs", "logp", "conf", "deltas"))
    probs, logp, conf, deltas, boxes = orc.resolve(pred, a64, shp.input_hw, shp.num_classes, log_softmax=True)
    ids, scores = orc.score_argmax(probs, conf)
    np.testing.assert_allclose(out["probs"].cpu().numpy(), probs, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(out["logp"].cpu().numpy(), logp, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(out["conf"].cpu().numpy(), conf, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(out["boxes"].cpu().numpy(), boxes, rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(out["scores"].cpu().numpy(), scores, rtol=RTOL, atol=1e-7)
    assert np.array_equal(out["deltas"].cpu().numpy(), deltas)
    got_ids = out["class_ids"].cpu().numpy()
    assert got_ids.dtype == np.int64
    mism = got_ids != ids
    if mism.any():  # only legal where the two best class scores are within rounding of each other
        s = (probs * conf)[mism]
        top2 = np.sort(s, axis=-1)[:, -2:]
        assert np.all(top2[:, 1] - top2[:, 0] <= 1e-6 * top2[:, 1])
    assert mism.mean() < 1e-4


def _assert_dpred(got, ref):
    """dpred against the reference's autograd (golden) / the oracle: rtol 1e-4 above an absolute floor of
    1e-6 of the largest gradient (entries that are differences of nearly equal terms, e.g. p_k - 1 of a confident class,
    carry the fp32 rounding of BOTH sides; a float64 yardstick of the reference's own graph puts its fp32 autograd at
    7e-5 relative, oracle/gen_golden.py:gen_loss)."""
    scale = float(np.abs(ref[np.isfinite(ref)]).max()) if np.isfinite(ref).any() else 1.0
    with np.errstate(invalid="ignore"):
        rel = np.abs(got - ref) / (np.abs(ref) + 1e-6 * scale)
    print(f"dpred: max relative error {np.nanmax(rel):.2e} (floor 1e-6 of max |dpred| = {scale:.3e})")
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-6 * scale)


@pytest.mark.parametrize("name", ["tiny_160x96", "kitti_1248x384"])
def test_decode_vs_reference_golden(ops, golden, name):
    g = golden("decode_filter_" + name)
    shp = SHAPES[name]
    a64, a32 = anchors_dev(shp)
    pred = synth.clustered_pred(shp, int(g["batch"]), int(g["seed"]), anchors=a64)
    out = ops.decode_scores(dev(pred), a32, shp.input_hw, shp.num_classes,
                            ("class_ids", "scores", "boxes", "probs", "logp", "conf"))
    np.testing.assert_allclose(out["probs"].cpu().numpy(), g["probs"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(out["logp"].cpu().numpy(), g["logp"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(out["conf"].cpu().numpy(), g["conf"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(out["boxes"].cpu().numpy(), g["boxes"], rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(out["scores"].cpu().numpy(), g["scores"], rtol=RTOL, atol=1e-7)
    assert (out["class_ids"].cpu().numpy() != g["class_ids"]).mean() < 1e-4


def test_decode_empty_batch_and_bad_args(ops):
    from squeezedet_pytorch_b200._lib import SqdError
    shp = synth.TINY
    _, a32 = anchors_dev(shp)
    out = ops.decode_scores(torch.empty((0, shp.num_anchors, 8), device="cuda"), a32, shp.input_hw, 3)
    assert out["scores"].shape == (0, shp.num_anchors)
    with pytest.raises(SqdError):
        ops.decode_scores(torch.zeros((1, shp.num_anchors, 9), device="cuda"), a32, shp.input_hw, 3)
    with pytest.raises(SqdError):
        ops.decode_scores(torch.zeros((1, shp.num_anchors, 8)), a32, shp.input_hw, 3)  # CPU tensor: no fallback


# ------------------------------------------------------------------------------------------------------
# a8-a9 filter
# ------------------------------------------------------------------------------------------------------
def _check_rows(rows, expect, exact_values):
    for b, (row, exp) in enumerate(zip(rows, expect)):
        n = len(exp["anchor_idx"])
        if n == 0:
            assert row is None
            continue
        assert row is not None, f"image {b}: nothing kept, expected {n}"
        assert np.array_equal(row["anchor_idx"].numpy(), exp["anchor_idx"]), f"image {b}"
        assert np.array_equal(row["class_ids"].numpy(), exp["class_ids"])
        if exact_values:
            assert np.array_equal(row["scores"].numpy(), exp["scores"])
            assert np.array_equal(row["boxes"].numpy(), exp["boxes"])
        else:
            np.testing.assert_allclose(row["scores"].numpy(), exp["scores"], rtol=RTOL, atol=1e-7)
            np.testing.assert_allclose(row["boxes"].numpy(), exp["boxes"], rtol=RTOL, atol=1e-3)


@pytest.mark.parametrize("name", ["tiny_160x96", "kitti_1248x384"])
def test_filter_on_reference_dense_outputs_bit_exact(ops, golden, name):
    """The reference's own dense (ids, scores, boxes) in -> the reference's kept set out, bit for bit."""
    g = golden("decode_filter_" + name)
    shp = SHAPES[name]
    det = ops.topk_nms(dev(g["class_ids"], torch.int64), dev(g["scores"]), dev(g["boxes"]), shp.num_classes,
                       shp.top_k, shp.nms_thresh, shp.score_thresh)
    expect = [dict(anchor_idx=i, class_ids=c, scores=s, boxes=x) for i, c, s, x in zip(
        split_ragged(g["kept_count"], g["kept_anchor"]), split_ragged(g["kept_count"], g["kept_class"]),
        split_ragged(g["kept_count"], g["kept_score"]), split_ragged(g["kept_count"], g["kept_box"]))]
    _check_rows(dets_to_lists(det), expect, exact_values=True)
    # padding rows are deterministic
    cnt = det.count.cpu().numpy()
    assert np.array_equal(cnt, g["kept_count"])
    assert (det.anchor.cpu().numpy()[0, cnt[0]:] == -1).all()


@pytest.mark.parametrize("name,batch", [("tiny_160x96", 6), ("kitti_1248x384", 4), ("stress_2496x768", 2)])
def test_fused_detect_equals_unfused_and_oracle(ops, name, batch):
    shp = SHAPES[name]
    a64, a32 = anchors_dev(shp)
    pred = synth.clustered_pred(shp, batch, 202, anchors=a64)
    dp = dev(pred)
    fused = ops.detect_from_pred(dp, a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
    dense = ops.decode_scores(dp, a32, shp.input_hw, shp.num_classes)
    unfused = ops.topk_nms(dense["class_ids"], dense["scores"], dense["boxes"], shp.num_classes, shp.top_k,
                           shp.nms_thresh, shp.score_thresh)
    # (i) the two CUDA routes agree bit for bit
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(fused, f), getattr(unfused, f)), f
    # (ii) the oracle's filter run on the CUDA dense outputs gives the identical kept set and values
    ids, sc, bx = (dense[k].cpu().numpy() for k in ("class_ids", "scores", "boxes"))
    expect = [orc.filter_image(ids[b], sc[b], bx[b], shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
              for b in range(batch)]
    _check_rows(dets_to_lists(fused), expect, exact_values=True)
    assert sum(len(e["anchor_idx"]) for e in expect) > 0
    # (iii) and the all-oracle route (numpy exp instead of CUDA expf) keeps the same anchors
    expect2 = orc.detect_filtered(pred, a64, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh,
                                  shp.score_thresh)
    _check_rows(dets_to_lists(fused), expect2, exact_values=False)


@pytest.mark.parametrize("name,batch", [("tiny_160x96", 5), ("kitti_1248x384", 9), ("stress_2496x768", 3)])
@pytest.mark.parametrize("score_thresh", [None, -1.0])
def test_detect_two_phase_equals_clustered(ops, name, batch, score_thresh):
    """sqd_detect_from_pred with scratch (streaming scan -> candidate lists -> per-image tail) and without (one
    clustered launch) are the same function; score_thresh -1 sends EVERY anchor through the candidate lists
    (list length == A, mid-scan compactions in the tail)."""
    shp = SHAPES[name]
    thr = shp.score_thresh if score_thresh is None else score_thresh
    a64, a32 = anchors_dev(shp)
    dp = dev(synth.clustered_pred(shp, batch, 404, anchors=a64))
    two = ops.detect_from_pred(dp, a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, thr, two_phase=True)
    one = ops.detect_from_pred(dp, a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, thr, two_phase=False)
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(two, f), getattr(one, f)), f
    assert int(two.count.sum()) > 0


@pytest.mark.parametrize("name,batch", [("tiny_160x96", 5), ("kitti_1248x384", 597), ("stress_2496x768", 3)])
@pytest.mark.parametrize("score_thresh", [None, -1.0])
def test_tail_cta_size_does_not_change_results(ops, name, batch, score_thresh):
    """The per-image tail runs in 512-thread CTAs for small batches and in 128-thread CTAs (8 per SM) above 592 images
    (KITTI x 597 takes the small CTA by default); SQD_TAIL_THREADS forces either.  Same results, including the lists of
    A candidates that score_thresh -1 produces (histogram select from L2 re-reads, running-threshold fallback)."""
    shp = SHAPES[name]
    thr = shp.score_thresh if score_thresh is None else score_thresh
    if batch > 100 and score_thresh is not None:
        batch = 40
    a64, a32 = anchors_dev(shp)
    base = synth.clustered_pred(shp, min(batch, 16), 407, anchors=a64)
    dp = dev(base).repeat((batch + 15) // 16, 1, 1)[:batch].contiguous()
    args = (dp, a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, thr)
    default = ops.detect_from_pred(*args, two_phase=True)
    with _lib.option("SQD_TAIL_THREADS", 128):
        small = ops.detect_from_pred(*args, two_phase=True)
    with _lib.option("SQD_TAIL_THREADS", 512):
        big = ops.detect_from_pred(*args, two_phase=True)
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(small, f), getattr(big, f)), f
        assert torch.equal(getattr(default, f), getattr(big, f)), f
    assert int(big.count.sum()) > 0


def test_detect_two_phase_generic_class_count(ops):
    """A class count without a specialised kernel (C = 5) goes through the generic staged scan."""
    shp = synth.Shape("c5", (96, 160), 5, 16)
    a64, a32 = anchors_dev(shp)
    rs = np.random.RandomState(7)
    pred = rs.standard_normal((3, shp.num_anchors, shp.num_fields)).astype(np.float32)
    dp = dev(pred)
    two = ops.detect_from_pred(dp, a32, shp.input_hw, 5, shp.top_k, shp.nms_thresh, 0.05, two_phase=True)
    one = ops.detect_from_pred(dp, a32, shp.input_hw, 5, shp.top_k, shp.nms_thresh, 0.05, two_phase=False)
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(two, f), getattr(one, f)), f
    expect = orc.detect_filtered(pred, a64, shp.input_hw, 5, shp.top_k, shp.nms_thresh, 0.05)
    _check_rows(dets_to_lists(two), expect, exact_values=False)


@pytest.mark.parametrize("name", list(SHAPES))
def test_fused_detect_vs_reference_golden(ops, golden, name):
    g = golden("decode_filter_" + name)
    shp = SHAPES[name]
    a64, a32 = anchors_dev(shp)
    pred = synth.clustered_pred(shp, int(g["batch"]), int(g["seed"]), anchors=a64)
    det = ops.detect_from_pred(dev(pred), a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh,
                               shp.score_thresh)
    expect = [dict(anchor_idx=i, class_ids=c, scores=s, boxes=x) for i, c, s, x in zip(
        split_ragged(g["kept_count"], g["kept_anchor"]), split_ragged(g["kept_count"], g["kept_class"]),
        split_ragged(g["kept_count"], g["kept_score"]), split_ragged(g["kept_count"], g["kept_box"]))]
    _check_rows(dets_to_lists(det), expect, exact_values=False)


def test_filter_edge_cases(ops):
    """Ties (lower anchor index first), nothing above threshold (None), fewer anchors than top_k,
    zero-area boxes (NaN IoU never suppresses), identical boxes."""
    A = 40
    ids = np.zeros((3, A), np.int64)
    scores = np.zeros((3, A), np.float32)
    boxes = np.zeros((3, A, 4), np.float32)
    # image 0: all scores tied at 0.5, all boxes disjoint -> top-8 = anchors 0..7 in index order
    scores[0] = 0.5
    boxes[0, :, 0] = np.arange(A) * 20
    boxes[0, :, 2] = boxes[0, :, 0] + 10
    boxes[0, :, 3] = 10
    # image 1: nothing above the score threshold
    scores[1] = 0.1
    boxes[1] = boxes[0]
    # image 2: identical boxes in two classes + zero-area boxes
    scores[2] = np.linspace(0.9, 0.4, A)
    boxes[2, :, :] = [5, 5, 50, 50]
    ids[2, 1::2] = 1
    boxes[2, 10:14] = 0.0
    det = ops.topk_nms(dev(ids), dev(scores), dev(boxes), 2, 8, 0.4, 0.3)
    rows = det.to_list()
    assert rows[0]["anchor_idx"].tolist() == list(range(8))
    assert rows[1] is None
    for b in range(3):
        exp = orc.filter_image(ids[b], scores[b], boxes[b], 2, 8, 0.4, 0.3)
        if len(exp["anchor_idx"]) == 0:
            assert rows[b] is None
        else:
            assert rows[b]["anchor_idx"].tolist() == exp["anchor_idx"].tolist()
    # fewer anchors than top_k
    det = ops.topk_nms(dev(ids[:, :5]), dev(scores[:, :5]), dev(np.ascontiguousarray(boxes[:, :5])), 2, 64, 0.4, 0.3)
    assert det.count.cpu().tolist()[0] == 5


def test_filter_properties_full_size(ops):
    """KITTI batch 20 (BASELINE configs[1]) and stress: sortedness, NMS invariant, idempotence."""
    for shp, batch in ((synth.KITTI, 20), (synth.STRESS, 4)):
        a64, a32 = anchors_dev(shp)
        pred = synth.clustered_pred(shp, batch, 303, anchors=a64)
        det = ops.detect_from_pred(dev(pred), a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh,
                                   shp.score_thresh)
        dense = ops.decode_scores(dev(pred), a32, shp.input_hw, shp.num_classes)
        sc_all = dense["scores"].cpu().numpy()
        for b, row in enumerate(det.to_list()):
            assert row is not None
            cls, sc, bx, idx = (row[k].numpy() for k in ("class_ids", "scores", "boxes", "anchor_idx"))
            assert np.all(np.diff(cls) >= 0)                                   # classes ascending
            for c in np.unique(cls):
                s = sc[cls == c]
                assert np.all(np.diff(s) <= 0)                                 # scores descending inside a class
                bb = bx[cls == c]
                for i in range(len(bb)):                                       # no kept pair overlaps > thresh
                    iou = orc.pair_iou(np.repeat(bb[i:i + 1], len(bb), 0), bb)
                    iou[i] = 0
                    assert np.nanmax(iou) <= shp.nms_thresh + 1e-6
            assert np.all(sc > np.float32(shp.score_thresh))
            kth = np.sort(sc_all[b])[-shp.top_k]
            assert np.all(sc >= kth)                                           # only top-k members survive
            assert np.array_equal(sc, sc_all[b][idx])
        # idempotence: filtering again with only the kept anchors present keeps exactly the same set
        keep_mask = torch.zeros_like(dense["scores"], dtype=torch.bool)
        for b, row in enumerate(det.to_list()):
            keep_mask[b, row["anchor_idx"].cuda()] = True
        sc2 = torch.where(keep_mask, dense["scores"], torch.zeros_like(dense["scores"]))
        det2 = ops.topk_nms(dense["class_ids"], sc2, dense["boxes"], shp.num_classes, shp.top_k, shp.nms_thresh,
                            shp.score_thresh)
        assert torch.equal(det2.count, det.count)
        for b, n in enumerate(det.count.cpu().tolist()):
            assert torch.equal(det2.anchor[b, :n], det.anchor[b, :n])


# ------------------------------------------------------------------------------------------------------
# a11-a13 matcher / targets
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(SHAPES))
def test_matcher_vs_reference_golden(name, golden):
    from squeezedet_pytorch_b200 import targets
    g = golden("matcher_" + name)
    shp = SHAPES[name]
    anchors = synth.anchor_table(shp)
    m = targets.AnchorMatcher(anchors, shp.num_classes)
    boxes_l = []
    for i in range(int(g["n_img"])):
        _, boxes = synth.gt_boxes(shp, int(g["seed0"]) + i)
        if i % 6 == 5:
            boxes = np.repeat(boxes[:1], 12, axis=0)
            boxes[:, 2] = boxes[:, 0] + 3.0
            boxes[:, 3] = boxes[:, 1] + 2.0
        boxes_l.append(boxes)
    gb, _, gc = m.pack(boxes_l)
    idx, deltas = m.match(gb, gc)
    idx, deltas = idx.cpu().numpy(), deltas.cpu().numpy()
    exp_idx = split_ragged(g["count"], g["anchor_idx"])
    exp_dl = split_ragged(g["count"], g["deltas"])
    for i, b in enumerate(boxes_l):
        n = len(b)
        assert np.array_equal(idx[i, :n], exp_idx[i]), f"image {i}"          # bit-exact matched anchors
        assert (idx[i, n:] == -1).all()
        np.testing.assert_allclose(deltas[i, :n], exp_dl[i], rtol=1e-6, atol=1e-7)
        o_d, o_i = orc.match_anchors(b, anchors)
        assert np.array_equal(o_i, idx[i, :n])


def test_matcher_distance_fallback_and_dropin_signature(golden):
    from squeezedet_pytorch_b200 import targets
    g = golden("matcher_fallback")
    deltas, idx = targets.compute_deltas(g["boxes"], g["anchors"])          # reference signature, numpy in/out
    assert idx.dtype == np.int32 and deltas.dtype == np.float32
    assert np.array_equal(idx, g["anchor_idx"])
    np.testing.assert_allclose(deltas, g["deltas"], rtol=1e-6, atol=1e-7)
    with pytest.raises(AssertionError):                                       # boxes.py:14-15
        targets.compute_deltas(np.array([[5, 5, 5, 9]], np.float32), g["anchors"])
    d0, i0 = targets.compute_deltas(np.zeros((0, 4), np.float32), g["anchors"])
    assert d0.shape == (0, 4) and i0.shape == (0,)


def test_dense_targets_vs_oracle():
    from squeezedet_pytorch_b200 import targets
    for shp in (synth.TINY, synth.KITTI):
        anchors = synth.anchor_table(shp)
        m = targets.AnchorMatcher(anchors, shp.num_classes)
        cls_l, box_l = zip(*[synth.gt_boxes(shp, 900 + i) for i in range(5)])
        gb, gcl, gc = m.pack(list(box_l), list(cls_l))
        gt = m.dense_targets(gb, gcl, gc).cpu().numpy()
        for i in range(5):
            exp = orc.dense_targets(cls_l[i], box_l[i], anchors, shp.num_classes)
            np.testing.assert_allclose(gt[i], exp, rtol=1e-6, atol=1e-7)
            assert np.array_equal(gt[i] != 0, exp != 0)
        one = targets.prepare_annotations(cls_l[0], box_l[0], anchors, shp.num_classes)
        np.testing.assert_allclose(one, orc.dense_targets(cls_l[0], box_l[0], anchors, shp.num_classes), rtol=1e-6,
                                   atol=1e-7)


# ------------------------------------------------------------------------------------------------------
# a14-a16 loss
# ------------------------------------------------------------------------------------------------------
def _loss_inputs(shp, batch, seed):
    anchors = synth.anchor_table(shp)
    pred = synth.clustered_pred(shp, batch, seed, anchors=anchors)
    gts = []
    for b in range(batch):
        cls, boxes = synth.gt_boxes(shp, 1000 * seed + b)
        gts.append(orc.dense_targets(cls, boxes, anchors, shp.num_classes))
    return anchors, pred, np.stack(gts)


@pytest.mark.parametrize("name", ["tiny_160x96", "kitti_1248x384"])
def test_loss_vs_reference_autograd_golden(ops, golden, name):
    g = golden("loss_" + name)
    shp = SHAPES[name]
    anchors, pred, gt = _loss_inputs(shp, int(g["batch"]), int(g["seed"]))
    B = pred.shape[0]
    grad = torch.full((B, 4), 1.0 / B, device="cuda")
    losses, dpred = ops.loss_fwd_bwd(dev(pred), dev(gt), dev(anchors.astype(np.float32)), shp.input_hw,
                                     shp.num_classes, (1.0, 3.75, 100.0, 6.0), grad_loss=grad)
    losses = losses.cpu().numpy()
    np.testing.assert_allclose(losses[:, 0], g["class_loss"], rtol=RTOL)
    np.testing.assert_allclose(losses[:, 1] + losses[:, 2], g["score_loss"], rtol=RTOL)
    np.testing.assert_allclose(losses[:, 3], g["bbox_loss"], rtol=RTOL)
    np.testing.assert_allclose(losses.sum(1), g["loss"], rtol=RTOL)
    ref = g["dpred"]
    _assert_dpred(dpred.cpu().numpy(), ref)


def test_loss_stress_shape_vs_oracle_and_module_autograd(ops):
    """C=8 (generic field widths) against the oracle, and the nn.Module surface end to end:
    Loss(cfg)(pred, gt) -> loss.mean().backward() fills pred.grad like the reference's autograd."""
    from squeezedet_pytorch_b200 import config, model
    shp = synth.Shape("mid", (192, 320), 8, 32)
    anchors, pred, gt = _loss_inputs(shp, 3, 7)
    exp = orc.loss_forward(pred, gt, anchors, shp.input_hw, shp.num_classes)
    exp_d = orc.loss_backward(pred, gt, anchors, shp.input_hw, shp.num_classes, np.full((3,), 1 / 3))
    cfg = config.make_config(shp)
    mod = model.Loss(cfg).cuda()
    p = dev(pred).requires_grad_(True)
    loss, stats = mod(p, dev(gt))
    loss.mean().backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), exp["loss"], rtol=RTOL)
    np.testing.assert_allclose(stats["class_loss"].detach().cpu().numpy(), exp["class_loss"], rtol=RTOL)
    np.testing.assert_allclose(stats["score_loss"].detach().cpu().numpy(), exp["score_loss"], rtol=RTOL)
    np.testing.assert_allclose(stats["bbox_loss"].detach().cpu().numpy(), exp["bbox_loss"], rtol=RTOL)
    _assert_dpred(p.grad.cpu().numpy(), exp_d)


def test_loss_zero_objects_nan_and_determinism(ops):
    shp = synth.TINY
    anchors = synth.anchor_table(shp)
    a32 = dev(anchors.astype(np.float32))
    pred = dev(synth.clustered_pred(shp, 1, 5, anchors=anchors))
    gt = torch.zeros((1, shp.num_anchors, shp.num_classes + 9), device="cuda")
    losses, dpred = ops.loss_fwd_bwd(pred, gt, a32, shp.input_hw, shp.num_classes, (1.0, 3.75, 100.0, 6.0))
    l = losses.cpu().numpy()[0]
    assert np.isnan(l[0]) and np.isnan(l[1]) and np.isnan(l[3]) and np.isfinite(l[2])
    assert torch.isnan(dpred).all()
    # run-to-run determinism of the fixed-order reductions
    _, pk, gk = _loss_inputs(synth.KITTI, 2, 9)
    a = dev(synth.anchor_table(synth.KITTI).astype(np.float32))
    r1 = ops.loss_fwd_bwd(dev(pk), dev(gk), a, synth.KITTI.input_hw, 3, (1.0, 3.75, 100.0, 6.0))
    r2 = ops.loss_fwd_bwd(dev(pk), dev(gk), a, synth.KITTI.input_hw, 3, (1.0, 3.75, 100.0, 6.0))
    assert torch.equal(r1[0], r2[0]) and torch.equal(r1[1], r2[1])


# ------------------------------------------------------------------------------------------------------
# 8(f).1 boxes_postprocess
# ------------------------------------------------------------------------------------------------------
def test_boxes_postprocess_vs_reference_golden(ops, golden):
    import json
    from squeezedet_pytorch_b200.detector import _meta_record
    g = golden("postprocess")
    for i in range(int(g["n"])):
        meta = json.loads(str(g[f"meta_{i}"]))
        coll = {k: [v] if not isinstance(v, list) else np.asarray([v]) for k, v in meta.items()}
        coll = {k: (np.asarray(v) if k != "flipped" else np.asarray(v)) for k, v in coll.items()}
        rec = _meta_record(coll, 1)
        boxes = g[f"in_{i}"]
        n = boxes.shape[0]
        det = ops.Detections(count=torch.tensor([n], dtype=torch.int32, device="cuda"),
                             anchor=torch.zeros((1, 8), dtype=torch.int32, device="cuda"),
                             cls=torch.zeros((1, 8), dtype=torch.int32, device="cuda"),
                             score=torch.zeros((1, 8), device="cuda"), box=torch.zeros((1, 8, 4), device="cuda"))
        det.box[0, :n] = dev(boxes)
        ops.boxes_postprocess_(det, dev(rec))
        np.testing.assert_allclose(det.box[0, :n].cpu().numpy(), g[f"out_{i}"], rtol=1e-6, atol=1e-4)


# ------------------------------------------------------------------------------------------------------
# 8(f): result packing (+ KITTI text) and input pre-processing
# ------------------------------------------------------------------------------------------------------
def test_pack_results_and_kitti_text(ops):
    """Detections -> one packed (B,k,6) array (+ boxes_postprocess on the device) -> the reference's per-image dicts and
    the KITTI result text, against the oracle's filter / postprocess / formatter."""
    from squeezedet_pytorch_b200 import results
    shp = synth.KITTI
    a64, a32 = anchors_dev(shp)
    batch = 5
    pred = synth.clustered_pred(shp, batch, 909, anchors=a64)
    det = ops.detect_from_pred(dev(pred), a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
    metas = [{"orig_size": np.array([375, 1242, 3], np.int32), "scales": np.array([384 / 375., 1248 / 1242.], np.float32)}
             for _ in range(batch)]
    metas[2]["flipped"] = True
    metas[3]["drifts"] = np.array([-7, 11], np.int32)
    rec = np.zeros((batch, 10), np.float32)
    rec[:, 0:2] = [m["scales"] for m in metas]
    rec[2, 6] = 1242
    rec[3, 7:9] = [-7, 11]
    box_before = det.box.clone()
    packed, count = results.to_host(det, dev(rec))
    assert torch.equal(det.box, box_before)                      # inputs untouched
    expect = orc.detect_filtered(pred, a64, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
    rows = results.unpack(packed, count, metas)
    texts = results.kitti_texts(packed, count)
    for b, (row, exp) in enumerate(zip(rows, expect)):
        n = len(exp["anchor_idx"])
        assert int(count[b]) == n
        if n == 0:
            assert set(row) == {"image_meta"} and texts[b] == ""
            continue
        assert np.array_equal(row["class_ids"], exp["class_ids"])
        np.testing.assert_allclose(row["scores"], exp["scores"], rtol=RTOL, atol=1e-7)
        want = orc.boxes_postprocess(exp["boxes"], metas[b])
        np.testing.assert_allclose(row["boxes"], want, rtol=RTOL, atol=2e-3)
        # the text is a pure function of the packed values
        assert texts[b] == orc.kitti_result_text(row["class_ids"], row["scores"], row["boxes"], results.KITTI_CLASS_NAMES)
        assert (packed[b, n:, 0] == -1).all()


def test_preprocess_vs_reference_golden(ops, golden):
    """whiten + cv2 bilinear resize + HWC->CHW on the GPU against BaseDataset.preprocess run unmodified (uint8 and
    float32 inputs, down/up-scaling, identity size)."""
    g = golden("preprocess")
    for i in range(int(g["n"])):
        seed, h0, w0, h, w = (int(v) for v in g[f"case_{i}"])
        img = np.random.RandomState(seed).randint(0, 256, size=(h0, w0, 3)).astype(np.uint8)
        for arr in (img, img.astype(np.float32)):
            out = ops.preprocess_images(dev(arr[None]), g["mean"], g["std"], (h, w))
            assert out.shape == (1, 3, h, w)
            np.testing.assert_allclose(out[0].cpu().numpy(), g[f"out_{i}"], rtol=1e-5, atol=6e-5)  # IPP fp32 coordinates, see test_oracle_golden
    # batched call == per-image calls, and the oracle agrees at the KITTI size (375x1242 -> 384x1248)
    big = np.random.RandomState(5).randint(0, 256, size=(2, 375, 1242, 3)).astype(np.uint8)
    out = ops.preprocess_images(dev(big), g["mean"], g["std"], synth.KITTI.input_hw).cpu().numpy()
    for b in range(2):
        np.testing.assert_allclose(out[b], orc.preprocess_image(big[b], g["mean"], g["std"], synth.KITTI.input_hw),
                                   rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------------
# filter: maximum sizes, tie policy under massive exact ties, odd shapes (all three CUDA routes + the oracle)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,k,A,levels", [(3, 64, 5000, 3), (1, 1, 777, 2), (8, 256, 9000, 5), (32, 1024, 20000, 4),
                                          (2, 7, 1031, 1), (20, 1024, 3000, 0)])
def test_filter_ties_and_max_sizes(ops, C, k, A, levels):
    """Logits quantised to a handful of values make thousands of anchors share EXACTLY the same score (levels = 1: every
    anchor ties), so the declared tie policy (lower anchor index first) decides the top-k, the histogram select of the
    two-phase tail overflows its buffer and falls back to the running-threshold radix select, and k = 1024 / C = 32
    exercise the size limits.  levels = 0: continuous scores.  All routes must agree bit for bit; the oracle's filter
    run on the CUDA dense outputs must give the same kept set."""
    rs = np.random.RandomState(1000 + C * 7 + k)
    B = 3
    shp = synth.Shape("x", (96, 160), C, k)
    pred = rs.standard_normal((B, A, C + 5)).astype(np.float32)
    if levels:
        pred[..., :C + 1] = np.round(pred[..., :C + 1] * levels / 2) * (2.0 / levels)
        if levels == 1:
            pred[..., :C + 1] = 0.5
    pred[..., C] += 1.0                                                     # enough confidence to pass a 0.05 threshold
    anchors = np.concatenate([rs.uniform(20, 140, (A, 1)), rs.uniform(20, 76, (A, 1)), rs.uniform(8, 60, (A, 2))], 1)
    a32 = dev(anchors.astype(np.float32))
    dp = dev(pred)
    thr, nms = 0.05, 0.4
    two = ops.detect_from_pred(dp, a32, shp.input_hw, C, k, nms, thr, two_phase=True)
    one = ops.detect_from_pred(dp, a32, shp.input_hw, C, k, nms, thr, two_phase=False)
    dense = ops.decode_scores(dp, a32, shp.input_hw, C)
    unf = ops.topk_nms(dense["class_ids"], dense["scores"], dense["boxes"], C, k, nms, thr)
    with _lib.option("SQD_TAIL_THREADS", 128):     # the 128-thread tail CTA large batches use
        small = ops.detect_from_pred(dp, a32, shp.input_hw, C, k, nms, thr, two_phase=True)
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(two, f), getattr(one, f)), f
        assert torch.equal(getattr(two, f), getattr(unf, f)), f
        assert torch.equal(getattr(two, f), getattr(small, f)), f
    ids, sc, bx = (dense[x].cpu().numpy() for x in ("class_ids", "scores", "boxes"))
    expect = [orc.filter_image(ids[b], sc[b], bx[b], C, k, nms, thr) for b in range(B)]
    _check_rows(dets_to_lists(two), expect, exact_values=True)
    assert int(two.count.sum()) > 0


@pytest.mark.parametrize("name", ["tiny_160x96", "kitti_1248x384"])
def test_nonfinite_logits_vs_reference_golden(ops, golden, name):
    """NaN / inf class and confidence logits against what the REFERENCE did with them (tests/golden/nonfinite_*.npz):
    torch ranks NaN scores first, so such an anchor enters the top-k, may suppress neighbours in NMS and is dropped by
    the final `score > thresh`.  Every CUDA route must return the reference's kept rows."""
    g = golden("nonfinite_" + name)
    shp = SHAPES[name]
    a64, a32 = anchors_dev(shp)
    pred = synth.nonfinite_pred(shp, int(g["seed"]), anchors=a64)
    with np.errstate(invalid="ignore"):
        exp = orc.detect_filtered(pred, a64, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)
    cls = split_ragged(g["kept_count"], g["kept_class"])
    sc = split_ragged(g["kept_count"], g["kept_score"])
    bx = split_ragged(g["kept_count"], g["kept_box"])
    dp = dev(pred)
    dense = ops.decode_scores(dp, a32, shp.input_hw, shp.num_classes)
    routes = {
        "two_phase": ops.detect_from_pred(dp, a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh, two_phase=True),
        "clustered": ops.detect_from_pred(dp, a32, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh, two_phase=False),
        "dense": ops.topk_nms(dense["class_ids"], dense["scores"], dense["boxes"], shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh),
    }
    for rname, det in routes.items():
        rows = det.to_list()
        for i, row in enumerate(rows):
            n = int(g["kept_count"][i])
            if n == 0:
                assert row is None, (rname, i)
                continue
            assert row is not None, (rname, i)
            assert np.array_equal(row["class_ids"].numpy(), cls[i]), (rname, i)
            assert np.array_equal(row["anchor_idx"].numpy(), exp[i]["anchor_idx"]), (rname, i)
            np.testing.assert_allclose(row["scores"].numpy(), sc[i], rtol=RTOL, atol=1e-7)
            np.testing.assert_allclose(row["boxes"].numpy(), bx[i], rtol=RTOL, atol=1e-3)


def test_filter_nan_and_inf_inputs(ops):
    """NaN / inf in the DELTAS too (there the reference asserts, modules.py:18, so its behaviour is undefined): nothing
    may hang or get corrupted and all CUDA routes still agree with each other."""
    shp = synth.TINY
    a64, a32 = anchors_dev(shp)
    pred = synth.clustered_pred(shp, 4, 5, anchors=a64)
    pred[0, 5, :] = np.nan
    pred[1, 7, 3] = np.inf
    pred[2, 9, 0] = -np.inf
    pred[3, 11, 4:] = np.inf
    dp = dev(pred)
    two = ops.detect_from_pred(dp, a32, shp.input_hw, 3, shp.top_k, 0.4, 0.3, two_phase=True)
    one = ops.detect_from_pred(dp, a32, shp.input_hw, 3, shp.top_k, 0.4, 0.3, two_phase=False)
    assert torch.equal(two.count, one.count) and torch.equal(two.anchor, one.anchor)
    assert int(two.count.max()) <= shp.top_k


def test_matcher_batched_rounds_equal_sequential(monkeypatch):
    """The batched rounds (many boxes per pass, accepted up to the first clash) against the box-by-box kernel and the
    oracle, on inputs built to clash: duplicated boxes (same best anchor), shifted near-duplicates, boxes with no
    overlapping anchor in between (distance fallback), more boxes than a round holds."""
    from squeezedet_pytorch_b200 import targets
    shp = synth.KITTI
    anchors = synth.anchor_table(shp)
    m = targets.AnchorMatcher(anchors, shp.num_classes)
    rs = np.random.RandomState(77)
    boxes_l = []
    for i in range(6):
        _, bx = synth.gt_boxes(shp, 400 + i)
        bx = np.concatenate([bx, bx[:3], bx[:2] + np.float32(0.25)])            # exact and near duplicates -> clashes
        if i % 2 == 0:
            tiny = np.array([[3.0, 3.0, 3.5, 3.4], [1240.0, 380.0, 1240.4, 380.3]], np.float32)   # overlap nothing? (fallback)
            bx = np.concatenate([bx[:2], tiny, bx[2:]])
        if i == 5:
            bx = np.concatenate([bx] * 4)[:70]                                  # more than two rounds of 32
        boxes_l.append(np.ascontiguousarray(bx, dtype=np.float32))
    gb, _, gc = m.pack(boxes_l)
    idx_b, del_b = m.match(gb, gc)
    with _lib.option("SQD_MATCH_SEQUENTIAL", 1):
        idx_s, del_s = m.match(gb, gc)
    assert torch.equal(idx_b, idx_s) and torch.equal(del_b, del_s)
    for i, bx in enumerate(boxes_l):
        exp_del, exp_idx = orc.match_anchors(bx, anchors)
        assert np.array_equal(idx_b[i, :len(bx)].cpu().numpy(), exp_idx), i
        np.testing.assert_allclose(del_b[i, :len(bx)].cpu().numpy(), exp_del, rtol=1e-6, atol=1e-6)


def test_filter_randomised_differential(ops):
    """tools/fuzz_parity.py as a test: 60 random configurations (class count 1..32, top-k 1..1024, 1..30,000 anchors,
    thresholds 0..1, tied scores, clustered boxes): every CUDA route agrees bit for bit with every other and with the
    oracle's Detector.filter restatement run on the CUDA dense outputs.  (1,650 further cases were run once: all exact.)"""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(root, "tools", "fuzz_parity.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rs = np.random.RandomState(20260)
    kept = sum(fuzz.one_case(rs, torch.device("cuda"))["kept"] for _ in range(60))
    assert kept > 0


def test_matcher_randomised_differential():
    """Random annotation sets against the oracle's sequential matcher: 1..200 boxes per image (more boxes than a batched
    round holds), tiny boxes far from every anchor (distance fallback), exact duplicates and boxes sitting exactly on
    anchors (IoU ties -> lowest anchor index), crowded scenes where boxes compete for the same anchors."""
    from squeezedet_pytorch_b200 import targets
    rs = np.random.RandomState(777)
    for shp in (synth.TINY, synth.KITTI):
        anchors = synth.anchor_table(shp)
        a_xyxy = orc.anchors_xyxy_f64(anchors)
        m = targets.AnchorMatcher(anchors, shp.num_classes)
        H, W = shp.input_hw
        for trial in range(6):
            boxes_l = []
            for _ in range(4):
                n = int(rs.choice([1, 2, 7, 33, 64, 120, 200]))
                kind = rs.randint(0, 4)
                if kind == 0:      # generic boxes
                    cx, cy = rs.uniform(0, W, n), rs.uniform(0, H, n)
                    w, h = rs.uniform(2, W / 3, n), rs.uniform(2, H / 2, n)
                elif kind == 1:    # tiny boxes: many have IoU 0 with every anchor after the first few are taken
                    cx, cy = rs.uniform(0, W, n), rs.uniform(0, H, n)
                    w, h = rs.uniform(1.01, 2.5, n), rs.uniform(1.01, 2.5, n)
                elif kind == 2:    # crowded: all boxes around one point
                    cx, cy = rs.normal(W / 2, 10, n), rs.normal(H / 2, 6, n)
                    w, h = rs.uniform(20, 60, n), rs.uniform(20, 60, n)
                else:              # exactly on anchors, with duplicates
                    pick = rs.randint(0, len(anchors), n)
                    pick[n // 2:] = pick[:n - n // 2]
                    bx = a_xyxy[pick]
                    cx, cy = (bx[:, 0] + bx[:, 2]) / 2, (bx[:, 1] + bx[:, 3]) / 2
                    w, h = bx[:, 2] - bx[:, 0] + 1, bx[:, 3] - bx[:, 1] + 1
                b = np.stack([cx - (w - 1) / 2, cy - (h - 1) / 2, cx + (w - 1) / 2, cy + (h - 1) / 2], 1).astype(np.float32)
                b = b[(b[:, 2] > b[:, 0]) & (b[:, 3] > b[:, 1])]                    # boxes.py:14-15
                boxes_l.append(b if len(b) else np.array([[1, 1, 9, 9]], np.float32))
            gb, _, gc = m.pack(boxes_l)
            idx, deltas = m.match(gb, gc)
            idx, deltas = idx.cpu().numpy(), deltas.cpu().numpy()
            for i, b in enumerate(boxes_l):
                n = len(b)
                o_d, o_i = orc.match_anchors(b, anchors)
                assert np.array_equal(o_i, idx[i, :n]), (shp.name, trial, i)
                assert len(set(idx[i, :n].tolist())) == n                            # every box its own anchor
                np.testing.assert_allclose(deltas[i, :n], o_d, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("C", [1, 2, 3, 5, 8, 20])
def test_loss_randomised_shapes_vs_oracle(ops, C):
    """Loss forward + backward for class counts with and without specialised kernels, random loss weights, random
    per-image upstream gradients, images with 1 object and with many: the four per-image terms and dpred against the
    oracle (numpy float32 restatement of Loss.forward and of its autograd backward)."""
    rs = np.random.RandomState(40 + C)
    shp = synth.Shape("rnd", (int(rs.choice([96, 192])), int(rs.choice([160, 320]))), C, 16)
    anchors = synth.anchor_table(shp)
    B = 4
    pred = (rs.standard_normal((B, shp.num_anchors, C + 5)) * rs.uniform(0.3, 2.0)).astype(np.float32)
    gts = []
    for b in range(B):
        cls, boxes = synth.gt_boxes(shp, 5000 + 10 * C + b, lo=1, hi=1 if b == 0 else 20)
        gts.append(orc.dense_targets(cls % C, boxes, anchors, C))
    gt = np.stack(gts)
    weights = tuple(float(x) for x in rs.uniform(0.5, 100.0, 4))
    gl = rs.uniform(0.1, 2.0, B)
    exp = orc.loss_forward(pred, gt, anchors, shp.input_hw, C, weights)
    exp_d = orc.loss_backward(pred, gt, anchors, shp.input_hw, C, gl, weights)
    grad = torch.from_numpy(np.repeat(gl[:, None], 4, 1).astype(np.float32)).cuda()
    losses, dpred = ops.loss_fwd_bwd(dev(pred), dev(gt), dev(anchors.astype(np.float32)), shp.input_hw, C, weights, grad_loss=grad)
    losses = losses.cpu().numpy()
    np.testing.assert_allclose(losses[:, 0], exp["class_loss"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(losses[:, 1] + losses[:, 2], exp["score_loss"], rtol=RTOL)
    np.testing.assert_allclose(losses[:, 3], exp["bbox_loss"], rtol=RTOL)
    np.testing.assert_allclose(losses.sum(1), exp["loss"], rtol=RTOL)
    _assert_dpred(dpred.cpu().numpy(), exp_d)


# ------------------------------------------------------------------------------------------------------
# BASELINE configs[0]: the demo (demo.py:17-52) -- sample PNG -> preprocess -> backbone -> head -> detect, batch 1
# ------------------------------------------------------------------------------------------------------
def test_demo_samples_vs_reference_golden(ops, golden):
    """Two of the reference's sample images (data/samples/kitti/testing/image_2, stored in the fixture as uint8) through
    OUR pipeline on the GPU -- sqd_preprocess, the stock backbone (cuDNN fp32, TF32 off), sqd_head_detect_fused,
    sqd_boxes_postprocess behind Detector.detect -- against what the reference's own preprocess / SqueezeDet / Detector
    returned on the CPU with the same seeded weights (oracle/gen_golden.py:gen_demo).  The backbone is stock PyTorch on
    both sides but cuDNN and the CPU kernels round differently, so the comparison has two parts: (1) on OUR Fire11
    features the head is exact against the oracle (kept anchors bit-exact); (2) end to end the kept anchors, classes and
    order equal the reference's, scores within 1e-3, boxes within 0.05 px of the original-image coordinates."""
    from squeezedet_pytorch_b200 import config as sqd_config
    from squeezedet_pytorch_b200.detector import Detector
    from squeezedet_pytorch_b200.model import SqueezeDet
    g = golden("demo_kitti_samples")
    shp = synth.KITTI
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = sqd_config.kitti_config(device="cuda")
        model = SqueezeDet(cfg)
        model.load_state_dict(synth.demo_state_dict(model, shp, int(g["seed"])))
        det = Detector(model, cfg)
        det_graph = Detector(model, sqd_config.kitti_config(device="cuda", cuda_graph=True))
        a64 = synth.anchor_table(shp)
        for i in range(int(g["n"])):
            rgb = g[f"image_{i}"]
            h0, w0 = rgb.shape[:2]
            x = ops.preprocess_images(dev(rgb[None]), cfg.rgb_mean.reshape(3), cfg.rgb_std.reshape(3), shp.input_hw)
            scales = np.array([shp.input_hw[0] / h0, shp.input_hw[1] / w0], dtype=np.float32)
            np.testing.assert_allclose(scales, g[f"scales_{i}"], rtol=1e-6)
            meta = {"image_id": [str(g["ids"][i])], "orig_size": torch.tensor([[h0, w0, 3]], dtype=torch.int32),
                    "scales": torch.from_numpy(scales)[None]}
            with torch.no_grad():
                feat = model.base.features(x)
            fs = feat[0, ::37, ::5, ::7].cpu().numpy()
            np.testing.assert_allclose(fs, g[f"feat_sample_{i}"], rtol=2e-3, atol=2e-4 * float(g[f"feat_absmax_{i}"]))
            # (1) our head on our features == the oracle on the same features, exactly
            raw, _ = det.detect_batch({"image": x}, postprocess=False)
            w = model.base.convdet.weight.detach().cpu().numpy()
            b = model.base.convdet.bias.detach().cpu().numpy()
            pred = orc.convdet_forward(feat.cpu().numpy(), w, b, shp.num_anchors, shp.num_fields)
            exp = orc.detect_filtered(pred, a64, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh)[0]
            n = int(raw.count[0])
            assert np.array_equal(raw.anchor[0, :n].cpu().numpy(), exp["anchor_idx"])
            # (2) end to end against the reference's Detector.detect -- eager and as a replayed CUDA graph (cfg.cuda_graph)
            gres = det_graph.detect({"image": x, "image_meta": meta})[0]
            gres = det_graph.detect({"image": x, "image_meta": meta})[0]      # second call: a replay
            res = det.detect({"image": x, "image_meta": meta})[0]
            for f in ("class_ids", "scores", "boxes"):
                assert np.array_equal(res[f], gres[f]), f
            assert np.array_equal(res["class_ids"], g[f"class_ids_{i}"])
            assert np.array_equal(raw.anchor[0, :n].cpu().numpy(), g[f"anchor_{i}"])
            np.testing.assert_allclose(res["scores"], g[f"scores_{i}"], rtol=1e-3)
            np.testing.assert_allclose(res["boxes"], g[f"boxes_{i}"], rtol=0, atol=0.05)
            assert len(res["class_ids"]) > 10
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


@pytest.mark.parametrize("thresh", [0.5, 0.25, 1.0 / 3.0, 0.4, 0.2, 0.0, 0.75, float(np.float32(1.0 / 3.0)),
                                    float(np.nextafter(np.float32(0.5), np.float32(0))), 1.0, 2.0, -0.1])
def test_nms_threshold_boundary_is_exact(ops, thresh):
    """The tail decides `float(inter / union) > thresh` without the division (inter vs mid * union in double, see
    nms_and_emit).  Boxes on a small integer lattice make thousands of pairs whose IoU is EXACTLY a small rational --
    1/2, 1/3, 2/5, 1/4 ... -- so thresholds on, just below and just above such values, 0, 1, > 1 and negative ones must
    all reproduce the reference arithmetic (oracle.nms = torchvision's CPU rules, pinned by nms_torchvision.npz); rows
    with zero-area, inverted and infinite boxes take the division path."""
    rs = np.random.RandomState(int(abs(thresh) * 1000) + 5)
    B, A, k = 6, 400, 256
    xy = rs.randint(0, 6, size=(B, A, 2)).astype(np.float32)
    wh = rs.randint(1, 7, size=(B, A, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], axis=2)
    boxes[0, :30, 2:] = boxes[0, :30, :2]                       # zero-area boxes (0/0 -> NaN: never suppress)
    boxes[1, :20, [0, 2]] = boxes[1, :20, [2, 0]]               # inverted boxes (negative areas)
    boxes[2, :10, 2] = np.inf                                   # infinite extent
    boxes[3] *= np.float32(1e-20)                               # tiny areas: products in the denormal range
    boxes[4] *= np.float32(3e18)                                # huge areas: area sums overflow to inf
    scores = rs.permutation(B * A).reshape(B, A).astype(np.float32) / (B * A) * 0.5 + 0.4    # distinct
    ids = np.zeros((B, A), dtype=np.int64)
    out = ops.topk_nms(dev(ids), dev(scores), dev(boxes), 1, k, thresh, 0.05)
    for b in range(B):
        exp = orc.filter_image(ids[b], scores[b], boxes[b], 1, k, thresh, 0.05)
        n = int(out.count[b])
        assert np.array_equal(out.anchor[b, :n].cpu().numpy(), exp["anchor_idx"]), (thresh, b)
