"""Randomised differential test of the decode + top-k + NMS kernels against the oracle (numpy restatement of
Detector.filter / torchvision nms): random class counts, anchor counts, top-k, thresholds, score distributions with
exact ties, clustered boxes.  The oracle's filter runs on the CUDA dense outputs, so the comparison is bit exact.
usage: python tools/fuzz_parity.py [cases] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from squeezedet_pytorch_b200 import ops  # noqa: E402


def one_case(rs, dev):
    C = int(rs.choice([1, 2, 3, 3, 3, 5, 8, 8, 13, 20, 32]))
    k = int(rs.choice([1, 2, 7, 16, 64, 64, 100, 256, 1024]))
    A = int(rs.choice([1, 5, 40, 333, 1031, 4096, 16848, 30000]))
    B = int(rs.randint(1, 5))
    levels = int(rs.choice([0, 0, 1, 2, 4, 16]))
    thr = float(rs.choice([0.0, 0.05, 0.3, 0.3, 0.6, 0.95]))
    nms = float(rs.choice([0.0, 0.2, 0.4, 0.4, 0.7, 1.0]))
    H, W = int(rs.choice([96, 384, 768])), int(rs.choice([160, 1248, 2496]))
    pred = rs.standard_normal((B, A, C + 5)).astype(np.float32)
    if levels:
        pred[..., :C + 1] = np.round(pred[..., :C + 1] * levels / 2) * (2.0 / levels)
    pred[..., C] += rs.uniform(-3, 2)
    # clustered boxes: a few centres, small deltas -> NMS has work to do
    nc = int(rs.randint(1, 12))
    centres = np.stack([rs.uniform(0, W, nc), rs.uniform(0, H, nc)], 1)
    which = rs.randint(0, nc, A)
    anchors = np.concatenate([centres[which] + rs.normal(0, 6, (A, 2)), rs.uniform(4, min(H, W) / 2, (A, 2))], 1)
    if rs.rand() < 0.3:
        pred[..., C + 1:] *= 0.05
    a32 = torch.from_numpy(anchors.astype(np.float32)).to(dev)
    dp = torch.from_numpy(pred).to(dev)
    two = ops.detect_from_pred(dp, a32, (H, W), C, k, nms, thr, two_phase=True)
    one = ops.detect_from_pred(dp, a32, (H, W), C, k, nms, thr, two_phase=False)
    dense = ops.decode_scores(dp, a32, (H, W), C)
    unf = ops.topk_nms(dense["class_ids"], dense["scores"], dense["boxes"], C, k, nms, thr)
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(two, f), getattr(one, f)), ("two-phase vs clustered", f)
        assert torch.equal(getattr(two, f), getattr(unf, f)), ("fused vs unfused", f)
    ids, sc, bx = (dense[x].cpu().numpy() for x in ("class_ids", "scores", "boxes"))
    rows = two.to_list()
    kept = 0
    for b in range(B):
        exp = orc.filter_image(ids[b], sc[b], bx[b], C, k, nms, thr)
        n = len(exp["anchor_idx"])
        kept += n
        if n == 0:
            assert rows[b] is None, ("oracle keeps nothing", b)
            continue
        assert rows[b] is not None, ("oracle keeps", n, b)
        assert np.array_equal(rows[b]["anchor_idx"].numpy(), exp["anchor_idx"]), ("kept anchors", b)
        assert np.array_equal(rows[b]["class_ids"].numpy(), exp["class_ids"]), ("classes", b)
        assert np.array_equal(rows[b]["scores"].numpy(), exp["scores"]), ("scores", b)
        assert np.array_equal(rows[b]["boxes"].numpy(), exp["boxes"]), ("boxes", b)
    return dict(C=C, k=k, A=A, B=B, levels=levels, thr=thr, nms=nms, kept=kept)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda")
    rs = np.random.RandomState(seed)
    total_kept = 0
    for i in range(cases):
        state = rs.get_state()
        try:
            info = one_case(rs, dev)
        except AssertionError as e:
            rs.set_state(state)
            print("CASE %d FAILED: %s" % (i, e))
            raise
        total_kept += info["kept"]
    print("fuzz: %d cases exact (kept %d detections in total), seed %d" % (cases, total_kept, seed))


if __name__ == "__main__":
    main()
