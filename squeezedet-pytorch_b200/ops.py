"""Tensor-level entry points: one Python function per C-ABI call (include/sqdet_b200.h).

Each function takes / returns CUDA torch tensors and enqueues work on torch's current stream.
Nothing here computes on the host; nothing falls back to PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import CONV_SIMT_FP32, CONV_TCGEN05_F16X3, LAYOUT_NCHW, LAYOUT_NHWC, check, load, ptr, stream_ptr, workspace


@dataclass
class Detections:
    """Batched filter output: rows [0, count[b]) of each (B, top_k[,4]) buffer are valid, in the
    reference's order (class ascending, score descending inside a class)."""
    count: torch.Tensor    # (B,) int32
    anchor: torch.Tensor   # (B, top_k) int32 kept anchor indices, -1 padding
    cls: torch.Tensor      # (B, top_k) int32
    score: torch.Tensor    # (B, top_k) float32
    box: torch.Tensor      # (B, top_k, 4) float32 xyxy
    status: torch.Tensor = None   # (1,) int32 view of the tcgen05 pipeline status word of the call that filled this

    def check_status(self):
        """Raise if the tcgen05 pipeline of the producing call hit a bounded-wait timeout (preemption, time slicing,
        a debugger ...) and drained with incomplete results.  Reads one int; called from to_list(), where the host
        synchronises anyway."""
        if self.status is not None:
            code = int(self.status.cpu()[0])
            if code != 0:
                raise _lib.SqdError(f"tcgen05 ConvDet pipeline timed out (role {code}): the detections are incomplete")

    def to_list(self):
        """One host transfer for the whole batch -> list of reference-style dicts (or None when
        an image keeps nothing, like Detector.filter, detector.py:115-116)."""
        self.check_status()
        count = self.count.cpu().tolist()
        anchor, cls, score, box = self.anchor.cpu(), self.cls.cpu(), self.score.cpu(), self.box.cpu()
        out = []
        for b, n in enumerate(count):
            if n == 0:
                out.append(None)
                continue
            out.append({"class_ids": cls[b, :n].to(torch.int64), "scores": score[b, :n], "boxes": box[b, :n],
                        "anchor_idx": anchor[b, :n].to(torch.int64)})
        return out


def _alloc_detections(batch, top_k, device) -> Detections:
    return Detections(
        count=torch.empty((batch,), dtype=torch.int32, device=device),
        anchor=torch.empty((batch, top_k), dtype=torch.int32, device=device),
        cls=torch.empty((batch, top_k), dtype=torch.int32, device=device),
        score=torch.empty((batch, top_k), dtype=torch.float32, device=device),
        box=torch.empty((batch, top_k, 4), dtype=torch.float32, device=device),
    )


def feature_layout(feat: torch.Tensor):
    """(layout code, tensor to hand to the kernel) for a logical (B,C,H,W) feature map: a
    channels_last tensor is consumed in place (zero copy); anything else is made NCHW-contiguous."""
    if feat.dim() != 4:
        raise _lib.SqdError("features must be (B, C, H, W)")
    if feat.is_contiguous(memory_format=torch.channels_last) and not feat.is_contiguous():
        return LAYOUT_NHWC, feat
    return LAYOUT_NCHW, feat.contiguous()


def resolve_conv_algo(algo, cin, cout):
    """The tcgen05 kernels take Cin % 64 == 0 and Cout <= 128; any other head runs on the library's own fp32 CUDA-core
    implicit GEMM (Cin % 16 == 0, Cout <= 128).  Beyond that there is nothing to run it on (no library fallback)."""
    if algo != CONV_SIMT_FP32 and (cin % 64 != 0 or cout > 128):
        algo = CONV_SIMT_FP32
    if algo == CONV_SIMT_FP32 and (cin % 16 != 0 or cout > 128):
        raise _lib.SqdError(f"ConvDet head with Cin={cin}, Cout={cout} is outside the kernels' limits "
                            "(Cin % 16 == 0 and Cout <= 128)")
    return algo


# ---- a1 --------------------------------------------------------------------------------------------
def pack_convdet_weights(weight: torch.Tensor) -> torch.Tensor:
    """Derive the tcgen05 kernel's weight planes from base.convdet.weight (Cout,Cin,3,3)."""
    lib = load()
    w = weight.detach().contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    nbytes = lib.sqd_convdet_packed_weight_bytes(cout, cin)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    check(lib.sqd_convdet_pack_weights(ptr(w), cout, cin, ptr(packed), stream_ptr(w.device)), "sqd_convdet_pack_weights")
    return packed


def convdet_forward(feat, weight, bias, packed=None, algo=CONV_TCGEN05_F16X3, num_fields=None, check_status=False):
    """feat (B,Cin,gh,gw) -> pred (B, gh*gw*K, C+5) [or (B,gh,gw,Cout) when num_fields is None]."""
    lib = load()
    layout, x = feature_layout(feat)
    B, cin, gh, gw = x.shape
    w = weight.detach().contiguous()
    b = bias.detach().contiguous()
    cout = w.shape[0]
    algo = resolve_conv_algo(algo, cin, cout)
    if algo != CONV_SIMT_FP32 and packed is None:
        packed = pack_convdet_weights(w)
    nbytes = lib.sqd_convdet_workspace_bytes(B, cin, gh, gw, cout, layout, algo)
    ws = workspace().get("convdet", nbytes, x.device)
    pred = torch.empty((B, gh, gw, cout), dtype=torch.float32, device=x.device)
    xp = C.c_void_p(x.data_ptr())  # channels_last tensors are not .is_contiguous(); pointer is still the base
    check(lib.sqd_convdet_forward(xp, layout, ptr(packed), ptr(w), ptr(b), B, cin, gh, gw, cout, ptr(pred), ptr(ws),
                                  ws.numel(), algo, stream_ptr(x.device)), "sqd_convdet_forward")
    if check_status and algo != CONV_SIMT_FP32:
        check(lib.sqd_convdet_status(ptr(ws), stream_ptr(x.device)), "sqd_convdet_status")
    if num_fields is not None:
        pred = pred.view(B, gh * gw * (cout // num_fields), num_fields)
    return pred


# ---- 8f.2 (first half): ConvDet backward w.r.t. features and bias ---------------------------------------
def pack_convdet_dgrad_weights(weight: torch.Tensor) -> torch.Tensor:
    """Flipped / transposed weight planes for convdet_dgrad; derived data, once per weight update."""
    lib = load()
    w = weight.detach().contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    packed = torch.empty(lib.sqd_convdet_dgrad_packed_bytes(cout, cin), dtype=torch.uint8, device=w.device)
    check(lib.sqd_convdet_dgrad_pack_weights(ptr(w), cout, cin, ptr(packed), stream_ptr(w.device)),
          "sqd_convdet_dgrad_pack_weights")
    return packed


def convdet_dgrad(gpred, weight, dgrad_packed=None):
    """gpred (B,gh,gw,Cout) [gradient of convdet_forward's output] -> gradient of the features, logical shape
    (B,Cin,gh,gw) in channels_last memory (a zero-copy permute of the kernel's NHWC output)."""
    lib = load()
    g = gpred.contiguous().float()
    B, gh, gw, cout = g.shape
    cin = weight.shape[1]
    if dgrad_packed is None:
        dgrad_packed = pack_convdet_dgrad_weights(weight)
    ws = workspace().get("convdet_dgrad", lib.sqd_convdet_dgrad_workspace_bytes(B, cin, gh, gw, cout), g.device)
    out = torch.empty((B, gh, gw, cin), dtype=torch.float32, device=g.device)
    check(lib.sqd_convdet_dgrad(ptr(g), ptr(dgrad_packed), B, cin, gh, gw, cout, ptr(out), ptr(ws), ws.numel(),
                                stream_ptr(g.device)), "sqd_convdet_dgrad")
    return out.permute(0, 3, 1, 2)


def convdet_wgrad(feat, gpred, tensor_cores=True, check_status=False):
    """feat (B,Cin,gh,gw) NCHW fp32, gpred (B,gh,gw,Cout) -> gradient of the ConvDet weight (Cout,Cin,3,3).
    tensor_cores: the tcgen05 f16x3 kernel where the shape allows (Cin % 64 == 0, Cout <= 80, even grid width);
    every other shape (e.g. the stress head's 117 channels) runs on the library's fp32 CUDA-core kernel."""
    lib = load()
    x = feat.detach().contiguous().float()
    g = gpred.contiguous().float()
    B, cin, gh, gw = x.shape
    cout = g.shape[-1]
    out = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=g.device)
    if tensor_cores and cin % 64 == 0 and cout <= 80 and gw % 2 == 0:     # the limits of sqd_convdet_wgrad_tc
        ws = workspace().get("convdet_wgrad_tc", lib.sqd_convdet_wgrad_tc_workspace_bytes(B, cin, gh, gw, cout), g.device)
        check(lib.sqd_convdet_wgrad_tc(ptr(x), ptr(g), B, cin, gh, gw, cout, ptr(out), ptr(ws), ws.numel(),
                                       stream_ptr(g.device)), "sqd_convdet_wgrad_tc")
        if check_status:
            check(lib.sqd_convdet_wgrad_tc_status(ptr(ws), stream_ptr(g.device)), "sqd_convdet_wgrad_tc_status")
        return out
    ws = workspace().get("convdet_wgrad", lib.sqd_convdet_wgrad_workspace_bytes(B, cin, gh, gw, cout), g.device)
    check(lib.sqd_convdet_wgrad(ptr(x), ptr(g), B, cin, gh, gw, cout, ptr(out), ptr(ws), ws.numel(), stream_ptr(g.device)),
          "sqd_convdet_wgrad")
    return out


def convdet_bias_grad(gpred):
    lib = load()
    g = gpred.contiguous().float()
    B, gh, gw, cout = g.shape
    gb = torch.empty((cout,), dtype=torch.float32, device=g.device)
    check(lib.sqd_convdet_bias_grad(ptr(g), B, gh, gw, cout, ptr(gb), stream_ptr(g.device)), "sqd_convdet_bias_grad")
    return gb


# ---- a2-a7 -----------------------------------------------------------------------------------------
def decode_scores(pred, anchors_f32, input_hw, num_classes, want=("class_ids", "scores", "boxes"), out=None):
    """pred (B,A,C+5) -> dict with the requested keys among class_ids/scores/boxes/probs/logp/conf/deltas.
    out: a dict returned by an earlier call with the same shapes, to be overwritten (no allocation)."""
    lib = load()
    pred = pred.contiguous()
    B, A, NF = pred.shape
    if NF != num_classes + 5:
        raise _lib.SqdError(f"pred has {NF} fields, expected {num_classes + 5}")
    dev = pred.device
    shapes = {"class_ids": ((B, A), torch.int64), "scores": ((B, A), torch.float32), "boxes": ((B, A, 4), torch.float32),
              "probs": ((B, A, num_classes), torch.float32), "logp": ((B, A, num_classes), torch.float32),
              "conf": ((B, A, 1), torch.float32), "deltas": ((B, A, 4), torch.float32)}
    if out is None:
        out = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=dev) for k in want}
    g = lambda k: ptr(out[k]) if k in out else None  # noqa: E731
    check(lib.sqd_decode_scores(ptr(pred), ptr(anchors_f32), B, A, num_classes, int(input_hw[0]), int(input_hw[1]),
                                g("class_ids"), g("scores"), g("boxes"), g("probs"), g("logp"), g("conf"), g("deltas"),
                                stream_ptr(dev)), "sqd_decode_scores")
    return out


# ---- a8-a9 -----------------------------------------------------------------------------------------
def topk_nms(class_ids, scores, boxes, num_classes, top_k, nms_thresh, score_thresh) -> Detections:
    """Dense SqueezeDet.forward outputs (B,A)/(B,A,4) -> Detections (Detector.filter for the batch)."""
    lib = load()
    class_ids, scores, boxes = class_ids.contiguous(), scores.contiguous(), boxes.contiguous()
    B, A = scores.shape
    det = _alloc_detections(B, top_k, scores.device)
    check(lib.sqd_topk_nms(ptr(class_ids), ptr(scores), ptr(boxes), B, A, num_classes, top_k, float(nms_thresh),
                           float(score_thresh), ptr(det.count), ptr(det.anchor), ptr(det.cls), ptr(det.score),
                           ptr(det.box), stream_ptr(scores.device)), "sqd_topk_nms")
    return det


def detect_from_pred(pred, anchors_f32, input_hw, num_classes, top_k, nms_thresh, score_thresh, two_phase=True,
                     out: Detections = None) -> Detections:
    """Fused decode + score + top-k + NMS from the ConvDet output.  two_phase: streaming scan into candidate lists
    + per-image tail (needs scratch); otherwise one clustered launch without scratch.  Same results."""
    lib = load()
    pred = pred.contiguous()
    B, A, _ = pred.shape
    det = out if out is not None else _alloc_detections(B, top_k, pred.device)
    ws, nbytes = None, 0
    if two_phase:
        nbytes = lib.sqd_detect_workspace_bytes(B, A)
        ws = workspace().get("detect", nbytes, pred.device)
    check(lib.sqd_detect_from_pred(ptr(pred), ptr(anchors_f32), B, A, num_classes, int(input_hw[0]), int(input_hw[1]),
                                   top_k, float(nms_thresh), float(score_thresh), ptr(det.count), ptr(det.anchor),
                                   ptr(det.cls), ptr(det.score), ptr(det.box), ptr(ws), nbytes, stream_ptr(pred.device)),
          "sqd_detect_from_pred")
    return det


def head_detect(feat, weight, bias, anchors_f32, anchors_per_grid, num_classes, input_hw, top_k, nms_thresh,
                score_thresh, packed=None, algo=CONV_TCGEN05_F16X3, out: Detections = None, slot=None) -> Detections:
    """Fire11 features -> final detections (ConvDet + decode + top-k + NMS) through one ABI call.

    `slot` (serving loops): calls issued on DIFFERENT streams must not share a workspace; give each stream its own slot
    number and its own `out`."""
    lib = load()
    layout, x = feature_layout(feat)
    B, cin, gh, gw = x.shape
    w = weight.detach().contiguous()
    b = bias.detach().contiguous()
    cout = w.shape[0]
    algo = resolve_conv_algo(algo, cin, cout)
    if algo != CONV_SIMT_FP32 and packed is None:
        packed = pack_convdet_weights(w)
    nbytes = lib.sqd_head_detect_workspace_bytes(B, cin, gh, gw, cout, layout, algo)
    ws = workspace().get("head_detect" if slot is None else "head_detect/%d" % slot, nbytes, x.device)
    det = out if out is not None else _alloc_detections(B, top_k, x.device)
    if algo != CONV_SIMT_FP32 and B > 0:
        off = lib.sqd_head_detect_status_offset(B, gh, gw, cout)
        det.status = ws[off:off + 4].view(torch.int32)
    check(lib.sqd_head_detect_fused(C.c_void_p(x.data_ptr()), layout, ptr(packed), ptr(w), ptr(b), ptr(anchors_f32), B,
                                    cin, gh, gw, anchors_per_grid, num_classes, int(input_hw[0]), int(input_hw[1]),
                                    top_k, float(nms_thresh), float(score_thresh), ptr(det.count), ptr(det.anchor),
                                    ptr(det.cls), ptr(det.score), ptr(det.box), ptr(ws), ws.numel(), algo,
                                    stream_ptr(x.device)), "sqd_head_detect_fused")
    return det


class HeadDetectLoop:
    """Device-resident serving loop over `slots` independent slots (stream + workspace + output block each).

    Consecutive batches are independent, so two of them may be in flight: the feature pre-pass of batch i+1 then runs on
    the SMs that the per-image tail of batch i leaves idle (batch 20: 20 of 148 SMs) -- +8-10 % images/s over issuing
    every call on one stream; three slots are slower (tools/two_slot_bench.py).  Every batch runs all of its kernels.

        loop = HeadDetectLoop(weight, bias, anchors, anchors_per_grid, num_classes, input_hw, top_k, nms, thr)
        for feat in batches:
            det = loop.submit(feat)      # enqueues; `det` is the slot's output block, valid once its stream has run
            ...                          # consume det on loop.stream_of(det) (or after loop.join()) before the slot
        loop.join()                      # comes round again `slots` submissions later

    `submit` makes the slot's stream wait for the caller's current stream (the features are ready there) and `join`
    makes the caller's stream wait for every slot."""

    def __init__(self, weight, bias, anchors_f32, anchors_per_grid, num_classes, input_hw, top_k, nms_thresh, score_thresh,
                 slots=2, packed=None, algo=CONV_TCGEN05_F16X3):
        self.args = (weight, bias, anchors_f32, anchors_per_grid, num_classes, input_hw, top_k, nms_thresh, score_thresh)
        self.packed, self.algo, self.top_k = packed, algo, top_k
        dev = weight.device
        self.device = dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, int(slots)))]
        self.dets = [None] * len(self.streams)
        self.i = 0
        if packed is None and resolve_conv_algo(algo, weight.shape[1], weight.shape[0]) != CONV_SIMT_FP32:
            self.packed = pack_convdet_weights(weight.detach().contiguous())

    def submit(self, feat) -> Detections:
        s = self.i % len(self.streams)
        self.i += 1
        st = self.streams[s]
        st.wait_stream(torch.cuda.current_stream(self.device))
        if self.dets[s] is None or self.dets[s].count.shape[0] != feat.shape[0]:
            self.dets[s] = _alloc_detections(feat.shape[0], self.top_k, self.device)
        with torch.cuda.stream(st):
            head_detect(feat, *self.args, packed=self.packed, algo=self.algo, out=self.dets[s], slot=s)
        feat.record_stream(st)
        return self.dets[s]

    def stream_of(self, det):
        return self.streams[next(i for i, d in enumerate(self.dets) if d is det)]

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)


def head_detect_profile(feat, bias, anchors_f32, anchors_per_grid, num_classes, input_hw, top_k, nms_thresh,
                        score_thresh, packed, out: Detections = None):
    """Diagnostic twin of head_detect: same kernels, returns (Detections, [split_ms, convdet_ms, filter_ms]) measured
    with CUDA events on the launching stream (sqd_head_detect_profile; synchronises)."""
    lib = load()
    layout, x = feature_layout(feat)
    B, cin, gh, gw = x.shape
    cout = anchors_per_grid * (num_classes + 5)
    nbytes = lib.sqd_head_detect_profile_workspace_bytes(B, cin, gh, gw, cout)
    ws = workspace().get("head_detect_profile", nbytes, x.device)
    det = out if out is not None else _alloc_detections(B, top_k, x.device)
    ms = (C.c_float * 3)()
    check(lib.sqd_head_detect_profile(C.c_void_p(x.data_ptr()), layout, ptr(packed), ptr(bias.detach().contiguous()),
                                      ptr(anchors_f32), B, cin, gh, gw, anchors_per_grid, num_classes, int(input_hw[0]),
                                      int(input_hw[1]), top_k, float(nms_thresh), float(score_thresh), ptr(det.count),
                                      ptr(det.anchor), ptr(det.cls), ptr(det.score), ptr(det.box), ptr(ws), ws.numel(),
                                      stream_ptr(x.device), ms), "sqd_head_detect_profile")
    return det, [float(v) for v in ms]


class HostDetections:
    """Pinned host buffers for head_detect_host (same fields as Detections)."""

    def __init__(self, batch, top_k):
        # ONE pinned block laid out like the device-side result block, the five fields are views of it: the call then
        # brings a batch's detections back with a single device-to-host copy (sqd_head_detect_host_result_layout)
        offs, total = (C.c_size_t * 5)(), C.c_size_t(0)
        check(load().sqd_head_detect_host_result_layout(int(batch), int(top_k), offs, C.byref(total)),
              "sqd_head_detect_host_result_layout")
        self.block = torch.empty((int(total.value),), dtype=torch.uint8).pin_memory()

        def view(i, n, dtype, shape):
            return self.block[offs[i]:offs[i] + n * 4].view(dtype).view(shape)
        self.count = view(0, batch, torch.int32, (batch,))
        self.anchor = view(1, batch * top_k, torch.int32, (batch, top_k))
        self.cls = view(2, batch * top_k, torch.int32, (batch, top_k))
        self.score = view(3, batch * top_k, torch.float32, (batch, top_k))
        self.box = view(4, batch * top_k * 4, torch.float32, (batch, top_k, 4))
        self.ready = None     # CUDA event of the call that last wrote these buffers (head_detect_host(sync=False))

    def wait(self):
        """Block until the call that fills these buffers has finished (no-op after a sync=True call)."""
        if self.ready is not None:
            self.ready.synchronize()
            self.ready = None
        return self

    def to_list(self):
        self.wait()
        return Detections(self.count, self.anchor, self.cls, self.score, self.box).to_list()


_copy_streams = {}


def head_detect_host(host_feat, weight, bias, anchors_f32, anchors_per_grid, num_classes, input_hw, top_k, nms_thresh,
                     score_thresh, packed=None, algo=CONV_TCGEN05_F16X3, out: HostDetections = None, chunk_images=5,
                     overlap=True, sync=True, slot=None) -> HostDetections:
    """HOST feature maps (B,Cin,gh,gw) fp32 (pinned, NCHW-contiguous) -> HOST detections through ONE ABI call that
    pipelines H2D copies with the kernels (sqd_head_detect_host).  `weight`/`bias`/`anchors` live on the device.

    Serving loop (keeps the PCIe link busy across calls): pass sync=False and alternate slot=0,1 with one
    HostDetections per slot; call `.wait()` on call i's result before issuing call i+2 (which reuses its slot).  Each
    slot has its own device workspace, so call i+1's copies start while call i's kernels still run."""
    lib = load()
    if host_feat.is_cuda or not host_feat.is_contiguous() or host_feat.dtype != torch.float32:
        raise _lib.SqdError("head_detect_host needs a contiguous fp32 HOST tensor (B,Cin,gh,gw)")
    dev = weight.device
    B, cin, gh, gw = host_feat.shape
    w = weight.detach().contiguous()
    b = bias.detach().contiguous()
    cout = w.shape[0]
    algo = resolve_conv_algo(algo, cin, cout)
    if algo != CONV_SIMT_FP32 and packed is None:
        packed = pack_convdet_weights(w)
    nbytes = lib.sqd_head_detect_host_workspace_bytes(B, cin, gh, gw, cout, top_k, LAYOUT_NCHW, algo, chunk_images)
    ws = workspace().get("head_detect_host" if slot is None else "head_detect_host/%d" % slot, nbytes, dev)
    det = out if out is not None else HostDetections(B, top_k)
    if det.ready is not None:
        raise _lib.SqdError("head_detect_host: the output buffers still belong to an unfinished call; .wait() on them first")
    flags = 0 if slot is None else 1     # SQD_HOST_NO_STAGING_FENCE: per-slot workspaces, reuse guarded by .wait()
    cst = None
    if overlap:
        key = torch.device(dev).index
        if key not in _copy_streams:
            _copy_streams[key] = torch.cuda.Stream(device=dev)
        cst = C.c_void_p(_copy_streams[key].cuda_stream)
    hp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    check(lib.sqd_head_detect_host(hp(host_feat), LAYOUT_NCHW, ptr(packed), ptr(w), ptr(b), ptr(anchors_f32), B, cin, gh,
                                   gw, anchors_per_grid, num_classes, int(input_hw[0]), int(input_hw[1]), top_k,
                                   float(nms_thresh), float(score_thresh), hp(det.count), hp(det.anchor), hp(det.cls),
                                   hp(det.score), hp(det.box), ptr(ws), ws.numel(), algo, int(chunk_images),
                                   stream_ptr(dev), cst, flags), "sqd_head_detect_host")
    if sync:
        torch.cuda.current_stream(dev).synchronize()
    else:
        det.ready = torch.cuda.Event()
        det.ready.record(torch.cuda.current_stream(dev))
    return det


# ---- a11-a13 ---------------------------------------------------------------------------------------
def match_anchors(gt_boxes, gt_count, anchors_f64):
    """gt_boxes (B,Gmax,4) f32 xyxy, gt_count (B,) i32, anchors (A,4) f64 -> (anchor_idx (B,Gmax) i32, deltas (B,Gmax,4))."""
    lib = load()
    gt_boxes, gt_count = gt_boxes.contiguous(), gt_count.contiguous()
    B, G, _ = gt_boxes.shape
    idx = torch.empty((B, G), dtype=torch.int32, device=gt_boxes.device)
    deltas = torch.empty((B, G, 4), dtype=torch.float32, device=gt_boxes.device)
    check(lib.sqd_match_anchors(ptr(gt_boxes), ptr(gt_count), B, G, ptr(anchors_f64), anchors_f64.shape[0], ptr(idx),
                                ptr(deltas), stream_ptr(gt_boxes.device)), "sqd_match_anchors")
    return idx, deltas


def build_targets(gt_boxes, gt_classes, gt_count, anchor_idx, deltas, num_anchors, num_classes):
    """-> dense gt (B, A, C+9) = [mask | box | deltas | one-hot] (datasets/base.py:61-76)."""
    lib = load()
    B, G, _ = gt_boxes.shape
    gt = torch.empty((B, num_anchors, num_classes + 9), dtype=torch.float32, device=gt_boxes.device)
    check(lib.sqd_build_targets(ptr(gt_boxes.contiguous()), ptr(gt_classes.contiguous()), ptr(gt_count.contiguous()),
                                ptr(anchor_idx), ptr(deltas), B, G, num_anchors, num_classes, ptr(gt),
                                stream_ptr(gt.device)), "sqd_build_targets")
    return gt


# ---- a14-a16 ---------------------------------------------------------------------------------------
def loss_fwd_bwd(pred, gt, anchors_f32, input_hw, num_classes, weights, grad_loss=None, want_grad=True):
    """-> (losses (B,4) = [class, positive_score, negative_score, bbox], dpred (B,A,C+5) or None)."""
    lib = load()
    pred, gt = pred.contiguous(), gt.contiguous()
    B, A, _ = pred.shape
    dev = pred.device
    nbytes = lib.sqd_loss_workspace_bytes(B, A)
    ws = workspace().get("loss", nbytes, dev)
    losses = torch.empty((B, 4), dtype=torch.float32, device=dev)
    dpred = torch.empty_like(pred) if want_grad else None
    w = (C.c_float * 4)(*[float(x) for x in weights])
    gl = grad_loss.contiguous().float() if grad_loss is not None else None
    check(lib.sqd_loss_fwd_bwd(ptr(pred), ptr(gt), ptr(anchors_f32), B, A, num_classes, int(input_hw[0]),
                               int(input_hw[1]), w, ptr(gl), ptr(losses), ptr(dpred), ptr(ws), ws.numel(),
                               stream_ptr(dev)), "sqd_loss_fwd_bwd")
    return losses, dpred


# ---- 8(f).1 ----------------------------------------------------------------------------------------
def boxes_postprocess_(det: Detections, meta: torch.Tensor):
    """In-place boxes_postprocess of the kept rows; meta (B,10) f32, see the header."""
    lib = load()
    B, K, _ = det.box.shape
    check(lib.sqd_boxes_postprocess(ptr(det.box), ptr(det.count), ptr(meta.contiguous()), B, K,
                                    stream_ptr(det.box.device)), "sqd_boxes_postprocess")
    return det


def pack_results(det: Detections, meta: torch.Tensor = None) -> torch.Tensor:
    """Detections (+ optional (B,10) postprocess records) -> (B, top_k, 6) [class, score, x1, y1, x2, y2]; det is
    not modified.  One .cpu() of this tensor (and of det.count) brings a whole batch to the host."""
    lib = load()
    B, K, _ = det.box.shape
    packed = torch.empty((B, K, 6), dtype=torch.float32, device=det.box.device)
    check(lib.sqd_pack_results(ptr(det.count), ptr(det.cls), ptr(det.score), ptr(det.box),
                               ptr(meta.contiguous()) if meta is not None else None, B, K, ptr(packed),
                               stream_ptr(det.box.device)), "sqd_pack_results")
    return packed


def format_kitti(packed_host: torch.Tensor, count_host: torch.Tensor, class_names):
    """HOST: packed (B,k,6) float32 + count (B,) int32 (CPU tensors) -> list of B strings, the content of the per-image
    result files KITTI.save_results writes (kitti.py:78-97)."""
    lib = load()
    if packed_host.is_cuda or count_host.is_cuda:
        raise _lib.SqdError("format_kitti formats HOST arrays (bring the packed results over with one .cpu())")
    packed_host = packed_host.contiguous().float()
    count_host = count_host.contiguous().to(torch.int32)
    B, K, _ = packed_host.shape
    names = (C.c_char_p * len(class_names))(*[n.lower().encode() for n in class_names])
    hp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    need = lib.sqd_format_kitti(hp(packed_host), hp(count_host), B, K, names, len(class_names), None, 0, None)
    if need < 0:
        check(int(need), "sqd_format_kitti")
    buf = C.create_string_buffer(int(need) + 1)
    offs = (C.c_longlong * (B + 1))()
    got = lib.sqd_format_kitti(hp(packed_host), hp(count_host), B, K, names, len(class_names), buf, int(need), offs)
    if got < 0:
        check(int(got), "sqd_format_kitti")
    raw = buf.raw
    return [raw[offs[b]:offs[b + 1]].decode() for b in range(B)]


def preprocess_images(images: torch.Tensor, mean, std, out_hw) -> torch.Tensor:
    """(B, H0, W0, 3) uint8 or float32 CUDA images (RGB, as loaded) -> (B, 3, H, W) float32 network input:
    whiten + bilinear resize + HWC->CHW (image.py whiten/resize, base.py:33)."""
    lib = load()
    if images.dim() != 4 or images.shape[-1] != 3 or images.dtype not in (torch.uint8, torch.float32):
        raise _lib.SqdError("preprocess_images needs (B, H, W, 3) uint8 or float32")
    images = images.contiguous()
    B, H0, W0, _ = images.shape
    out = torch.empty((B, 3, int(out_hw[0]), int(out_hw[1])), dtype=torch.float32, device=images.device)
    m = (C.c_float * 3)(*[float(x) for x in mean])
    s = (C.c_float * 3)(*[float(x) for x in std])
    check(lib.sqd_preprocess(ptr(images), 0 if images.dtype == torch.uint8 else 1, B, H0, W0, m, s, int(out_hw[0]),
                             int(out_hw[1]), ptr(out), stream_ptr(images.device)), "sqd_preprocess")
    return out
