#!/usr/bin/env python
"""Benchmark of the hot path: Fire11 features -> ConvDet -> decode -> top-k -> per-class NMS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one pass of the path over one batch of synthetic feature
maps (BASELINE.json configs[1]: KITTI 1248x384 eval shape, batch 20 per GPU, 78x24 grid, 9 anchors,
3 classes, top-64, NMS 0.4).  With N>1 every rank runs its own batch-20 slice (images are
independent: no collective on the inference path; weak scaling); time = max over ranks.

  value     images/s, inputs resident in HBM, K steps between two CUDA events, issued through ops.HeadDetectLoop
            (two slots: stream + workspace + output block each, steps alternate); `single_stream` = one stream
  e2e       images/s through the same public call with HOST (pinned) buffers: H2D of the step's
            features and D2H of its detections inside the timed region; bare_h2d_copy_floor beside it
  roofline  ConvDet tcgen05 kernel (the dominant launch): algorithmic FLOPs / event-timed duration; traffic parsed
            from the committed ncu summary under profiles/
  cpu_baseline / --impl reference: the reference's OWN files (oracle/_ref, copied by oracle/make_ref.py) on the
            host cores; the oracle port only if that copy did not travel
  config1_demo / config3 / train_step / config5_stress: the other BASELINE.json configurations
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "images/sec head+decode+NMS @1248x384 (whole job; per-GPU = value / n_gpus)"
FLOP_PER_IMAGE = 2 * 1872 * 72 * 6912          # 1,863,254,016
PRED_BYTES_PER_IMAGE = 16848 * 8 * 4           # 539,136
FEAT_BYTES_PER_IMAGE = 768 * 24 * 78 * 4       # 5,750,784
NCU_STEP_PROFILE = "profiles/r02_ncu_b20.txt"                  # ncu --set full of the four step kernels at B = 20
NCU_DECODE_PROFILE = "profiles/r02_ncu_decode_nms_b1024.txt"   # ... of scan / tail / dense decode on 1024 images
DENSE_BYTES_PER_IMAGE = 16848 * (8 + 4 + 16)   # 471,744 (int64 id, score, box)


def ncu_traffic(profile, *kernels):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the named kernels, read from a committed
    `tools/ncu_summary.py full` text summary of one `ncu --set full` capture; None if the file or a kernel is missing.
    (Nothing is profiled during the bench run: ncu cannot run inside the timed process.)"""
    path = os.path.join(ROOT, profile)
    if not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, cur, seen = 0.0, None, set()
    for ln in open(path):
        if ln.startswith("=="):
            cur = next((k for k in kernels if k in ln), None)
            if cur in seen:      # first launch of each kernel only
                cur = None
            elif cur:
                seen.add(cur)
        elif cur:
            parts = ln.split()
            if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[2] in unit:
                total += float(parts[1]) * unit[parts[2]]
    return int(total) if len(seen) == len(kernels) else None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt, self.ok = index, [], threading.Event(), False
        self.marks = {}
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, int(reasons)))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0, t1):
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
                 0x100: "display_clock_setting"}
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        reasons = 0
        for s in inside:
            reasons |= s[2]
        return {"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None,
                "sm_max_mhz": float(self.max_sm), "samples": len(inside),
                "reasons": sorted(n for bit, n in names.items() if reasons & bit)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference code itself cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_pass(orc, feat, w, b, anchors, shp):
    pred = orc.convdet_forward(feat, w, b, shp.num_anchors, shp.num_fields)
    return orc.detect_filtered(pred, anchors, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh,
                               shp.score_thresh)


def make_reference_runner(shp, w, b, device="cpu"):
    """The reference's OWN classes, imported from oracle/_ref (a verbatim copy made by oracle/make_ref.py; absent ->
    None): SqueezeDet with the backbone replaced by Identity, so SqueezeDetBase.forward's tail (squeezedet.py:79-87),
    PredictionResolver and SqueezeDet.forward (:109-120,197-206) run on Fire11-shaped input, then Detector.detect
    (detector.py:20-50: per-image filter with torchvision nms, boxes_postprocess) -- on the CPU, as cfg.device says."""
    import types
    import torch
    from oracle import make_ref
    ref = make_ref.load()
    if ref is None:
        return None
    from squeezedet_pytorch_b200 import synth
    anchors = ref.boxes.generate_anchors(shp.grid_hw, shp.input_hw, synth.KITTI_SEEDS)
    cfg = types.SimpleNamespace(
        input_size=shp.input_hw, num_classes=shp.num_classes, anchors=anchors, anchors_per_grid=shp.anchors_per_grid,
        num_anchors=anchors.shape[0], arch="squeezedet", dropout_prob=0.5, device=torch.device(device),
        keep_top_k=shp.top_k, nms_thresh=shp.nms_thresh, score_thresh=shp.score_thresh, debug=0, mode="eval",
        class_loss_weight=1.0, positive_score_loss_weight=3.75, negative_score_loss_weight=100.0, bbox_loss_weight=6.0)
    net = ref.model.SqueezeDet(cfg)
    net.base.features = torch.nn.Identity()
    with torch.no_grad():
        net.base.convdet.weight.copy_(torch.from_numpy(w))
        net.base.convdet.bias.copy_(torch.from_numpy(b))
    det = ref.detector.Detector(net, cfg)

    def run(feat_t):
        return det.detect({"image": feat_t, "image_meta": {}})
    return run


def run_cpu_arm(args, budget_s, warmup, steps):
    """Times the reference's CPU implementation of the path on all host cores: the reference's own files (oracle/_ref,
    kind "reference") when they travelled with the snapshot, else the oracle port (kind "port").
    Returns a dict describing the run."""
    import torch
    from squeezedet_pytorch_b200 import synth
    shp = synth.KITTI
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.batch
    feat = synth.features(shp, B, 1234)
    w, b = synth.convdet_params(shp, 4321)
    runner = None if args.cpu_port else make_reference_runner(shp, w, b)
    if runner is not None:
        feat_t = torch.from_numpy(feat)
        one = lambda: runner(feat_t)   # noqa: E731
        kind = "reference"
        what = ("the reference's own SqueezeDet (Identity backbone) + Detector.detect imported from oracle/_ref "
                f"(verbatim copy of src/model, src/engine/detector.py, src/utils), torch CPU {cores} threads")
    else:
        from oracle import oracle as orc
        anchors = synth.anchor_table(shp)
        one = lambda: cpu_reference_pass(orc, feat, w, b, anchors, shp)   # noqa: E731
        kind = "port"
        what = f"torch CPU conv2d ({cores} threads) + numpy decode + per-image top-k/NMS (oracle/oracle.py)"
    t0 = time.perf_counter()
    first = one()                                               # first pass: page-in / thread pool spin-up
    est = time.perf_counter() - t0
    assert len(first) == B
    warmup = max(1, min(warmup, int(max(1, 0.2 * budget_s / max(est, 1e-3)))))
    steps = max(1, min(steps, int(max(1, 0.8 * budget_s / max(est, 1e-3)))))
    for _ in range(warmup):
        one()
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        one()
        times.append(time.perf_counter() - t)
    per_step = float(np.mean(times))
    info = {"value": B / per_step, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": f"{steps} timed passes (after {warmup} warm-up) over one batch of {B} KITTI-shaped feature maps: " + what,
            "ms_per_step": per_step * 1e3, "steps_run": steps}
    return info


_RESULT_OUT = None


def emit_result(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    info = run_cpu_arm(args, budget_s=90.0, warmup=args.warmup, steps=args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": info["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "KITTI 1248x384 eval-shape head+decode+NMS, batch %d (BASELINE configs[1])" % args.batch,
                   "note": "reference's CPU implementation of the path on all host threads: kind 'reference' = the "
                           "reference's own files from oracle/_ref, kind 'port' = the oracle restatement (when the copy "
                           "did not travel)"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": info["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_result(line)
    return 0


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def demo_config(args, dev, ops, synth):
    """BASELINE configs[0] on the GPU (and, when oracle/_ref travelled, the reference's own demo pipeline on the CPU)."""
    import torch
    from squeezedet_pytorch_b200 import config as sqd_config
    from squeezedet_pytorch_b200.detector import Detector
    from squeezedet_pytorch_b200.model import SqueezeDet
    shp = synth.KITTI
    gpath = os.path.join(ROOT, "tests", "golden", "demo_kitti_samples.npz")
    if os.path.exists(gpath):
        g = np.load(gpath, allow_pickle=False)
        rgb, what = g["image_0"], "the reference's sample image %s (375x1242 uint8)" % str(g["ids"][0])
    else:
        rgb, what = np.random.RandomState(3).randint(0, 256, size=(375, 1242, 3)).astype(np.uint8), "synthetic 375x1242 uint8 image"
    h0, w0 = rgb.shape[:2]
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = sqd_config.kitti_config(device=dev)
        model = SqueezeDet(cfg)
        model.load_state_dict(synth.demo_state_dict(model, shp, 2024))
        det = Detector(model, cfg)
        host = torch.from_numpy(rgb).pin_memory()
        scales = torch.tensor([[shp.input_hw[0] / h0, shp.input_hw[1] / w0]], dtype=torch.float32)
        meta = {"image_id": ["demo"], "orig_size": torch.tensor([[h0, w0, 3]], dtype=torch.int32), "scales": scales}
        mean, std = cfg.rgb_mean.reshape(3), cfg.rgb_std.reshape(3)

        def one():
            x = ops.preprocess_images(host.to(dev, non_blocking=True)[None], mean, std, shp.input_hw)
            return det.detect({"image": x, "image_meta": meta})[0]
        for _ in range(5):
            res = one()
        n = 40
        t0 = time.perf_counter()
        for _ in range(n):
            res = one()                 # ends with the device-to-host copy of the detections: a blocking call
        ms_eager = (time.perf_counter() - t0) / n * 1e3
        # stage split with events (one more pass)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        stage = np.zeros(3)
        for _ in range(10):
            ev[0].record()
            x = ops.preprocess_images(host.to(dev, non_blocking=True)[None], mean, std, shp.input_hw)
            ev[1].record()
            with torch.no_grad():
                feat = model.base.features(x)
            ev[2].record()
            ops.head_detect(feat, model.base.convdet.weight, model.base.convdet.bias, model.resolver._anchors_on(feat.device),
                            cfg.anchors_per_grid, cfg.num_classes, cfg.input_size, cfg.keep_top_k, cfg.nms_thresh,
                            cfg.score_thresh, packed=model.base.packed_weights())
            ev[3].record()
            torch.cuda.synchronize()
            stage += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(3)]) / 10
        # the same call with cfg.cuda_graph: backbone + path captured once per input shape and replayed (the backbone turns
        # out to be GPU bound, not launch bound: fp32 cuDNN convolutions with TF32 off, so the replay gains ~5 %)
        det = Detector(model, sqd_config.kitti_config(device=dev, cuda_graph=True))
        for _ in range(5):
            gres = one()
        t0 = time.perf_counter()
        for _ in range(n):
            gres = one()
        ms = (time.perf_counter() - t0) / n * 1e3
        assert all(np.array_equal(res[f], gres[f]) for f in ("class_ids", "scores", "boxes"))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    out = {"workload": "BASELINE configs[0]: demo, batch 1: " + what + " on the host -> preprocess -> SqueezeDet backbone (stock "
                       "PyTorch / cuDNN fp32) -> ConvDet + decode + top-64 + NMS + boxes_postprocess -> result dict on the host; "
                       "seeded weights (the bundled checkpoint is not in the image)",
           "ms_per_image": ms, "images_per_s": 1e3 / ms, "kept": int(len(res.get("class_ids", []))),
           "mode": "Detector(cfg.cuda_graph=True): one graph replay per image", "eager_ms_per_image": ms_eager,
           "stage_ms": {"h2d_preprocess": float(stage[0]), "backbone_stock_pytorch": float(stage[1]),
                        "path_head_decode_nms": float(stage[2])}}
    if not args.no_cpu_baseline:
        out["reference_cpu"] = demo_reference_cpu(rgb, synth)
    return out


def demo_reference_cpu(rgb, synth):
    """The reference's own demo pipeline on the host cores for the same image: utils.image whiten / resize,
    model.squeezedet.SqueezeDet (backbone included) and engine.detector.Detector.detect, imported from oracle/_ref."""
    import types
    import torch
    from oracle import make_ref
    ref = make_ref.load()
    if ref is None:
        return {"unavailable": "oracle/_ref did not travel with the snapshot"}
    import importlib
    rimg = importlib.import_module("utils.image")
    shp = synth.KITTI
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    anchors = ref.boxes.generate_anchors(shp.grid_hw, shp.input_hw, synth.KITTI_SEEDS)
    cfg = types.SimpleNamespace(
        input_size=shp.input_hw, num_classes=shp.num_classes, anchors=anchors, anchors_per_grid=shp.anchors_per_grid,
        num_anchors=anchors.shape[0], arch="squeezedet", dropout_prob=0.5, device=torch.device("cpu"),
        keep_top_k=shp.top_k, nms_thresh=shp.nms_thresh, score_thresh=shp.score_thresh, debug=0, mode="eval",
        class_loss_weight=1.0, positive_score_loss_weight=3.75, negative_score_loss_weight=100.0, bbox_loss_weight=6.0)
    net = ref.model.SqueezeDet(cfg)
    net.load_state_dict(synth.demo_state_dict(net, shp, 2024))
    det = ref.detector.Detector(net, cfg)
    mean = np.array([93.877, 98.801, 95.923], dtype=np.float32).reshape(1, 1, 3)
    std = np.array([78.782, 80.130, 81.200], dtype=np.float32).reshape(1, 1, 3)

    def one():
        image = rgb.astype(np.float32)
        meta = {"image_id": "demo", "orig_size": np.array(image.shape, dtype=np.int32)}
        image, meta = rimg.whiten(image, meta, mean=mean, std=std)
        image, meta, _ = rimg.resize(image, meta, shp.input_hw)
        x = torch.from_numpy(image.transpose(2, 0, 1)).unsqueeze(0)
        bmeta = {k: torch.from_numpy(v).unsqueeze(0) if isinstance(v, np.ndarray) else [v] for k, v in meta.items()}
        return det.detect({"image": x, "image_meta": bmeta})[0]
    one()
    times = []
    for _ in range(5):
        t = time.perf_counter()
        res = one()
        times.append(time.perf_counter() - t)
    return {"ms_per_image": float(np.mean(times)) * 1e3, "images_per_s": 1.0 / float(np.mean(times)), "cores": cores,
            "kind": "reference", "kept": int(len(res.get("class_ids", []))),
            "sample": "5 timed images (after 1 warm-up), the reference's whiten/resize + SqueezeDet + Detector.detect on CPU"}


def main_ours(args):
    import torch
    import torch.distributed as dist
    from squeezedet_pytorch_b200 import _lib, ops, synth
    from squeezedet_pytorch_b200 import dist as sdist

    rank, world, local = sdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    shp = synth.KITTI
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    peaks = load_peaks()

    # ---- inputs (resident): R rotating feature buffers > L2 so every step reads HBM-cold features -------------
    R = 3
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats = [torch.relu(torch.randn((B, shp.in_channels, *shp.grid_hw), generator=gen, device=dev)) for _ in range(R)]
    if args.layout == "channels_last":
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    w_np, b_np = synth.convdet_params(shp, 4321)
    weight, bias = torch.from_numpy(w_np).to(dev), torch.from_numpy(b_np).to(dev)
    packed = ops.pack_convdet_weights(weight)
    anchors = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
    det = ops._alloc_detections(B, shp.top_k, dev)

    def step(i):
        return ops.head_detect(feats[i % R], weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw,
                               shp.top_k, shp.nms_thresh, shp.score_thresh, packed=packed, out=det)

    # Serving loop of the device-resident path: the steps alternate over S = 2 slots, each with its own stream, workspace
    # and output block, so that step i+1's feature pre-pass fills the SMs the per-image tail of step i leaves idle (at
    # batch 20 the tail runs on 20 of 148 SMs).  Every step does all of its work; only the order of independent steps'
    # kernels on the GPU changes (tools/two_slot_bench.py: 148 -> 137 us per step; three slots are slower).
    S = max(1, args.slots)
    loop = ops.HeadDetectLoop(weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw, shp.top_k,
                              shp.nms_thresh, shp.score_thresh, slots=S, packed=packed)

    def slot_step(i):
        return loop.submit(feats[i % R])

    def timed_slot_loop(n):
        """n steps through the serving loop between two events on the current stream (the slots fork from / join it)."""
        main = torch.cuda.current_stream(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        last = None
        for i in range(n):
            last = slot_step(i)
        loop.join()
        b.record(main)
        return a, b, last

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if sampler.ok:
        sampler.start()

    # ---- device-resident throughput --------------------------------------------------------------------------
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    barrier()
    # (a) one stream, K calls back to back: the latency-ordered chain (reported as `single_stream`)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_single = max_over_ranks(e0.elapsed_time(e1))
    # (b) the serving loop over S slots: `value`
    for i in range(max(W, 2 * S)):
        slot_step(i)
    torch.cuda.synchronize()
    barrier()
    t_wall0 = time.perf_counter()
    e0, e1, last_det = timed_slot_loop(K)
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    step(K - 1)   # the slot that ran step K-1 holds the same detections as the single-stream call
    torch.cuda.synchronize()
    for f in ("count", "anchor", "cls", "score", "box"):
        assert torch.equal(getattr(last_det, f), getattr(det, f)), f
    if sampler.ok and len([s for s in sampler.samples if t_wall0 <= s[0] <= t_wall1]) < 5:
        # timed region too short for NVML: keep the identical load running (untimed) until enough samples exist
        t_extra0 = time.perf_counter()
        i = 0
        while time.perf_counter() - t_extra0 < 1.5:
            slot_step(i)
            i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.summary(t_extra0 + 0.2, time.perf_counter())
        clocks["note"] = "timed region shorter than NVML sampling; sampled under the same load right after it"
    else:
        clocks = sampler.summary(t_wall0, t_wall1)
    check_count = det.count.cpu()
    assert int(check_count.min()) >= 0 and int(check_count.max()) <= shp.top_k
    value = world * B * K / (ms_total * 1e-3)

    # ---- per-kernel durations (rank 0): the same kernel sequence through sqd_head_detect_profile, which records
    #      CUDA events between the stages on the launching stream -------------------------------------------------
    kern = {}
    if rank == 0:
        nprof = min(K, 40)
        det_p = ops._alloc_detections(B, shp.top_k, dev)
        rows = []
        for i in range(nprof + 3):
            _, ms = ops.head_detect_profile(feats[i % R], bias, anchors, shp.anchors_per_grid, shp.num_classes,
                                            shp.input_hw, shp.top_k, shp.nms_thresh, shp.score_thresh, packed, out=det_p)
            if i >= 3:
                rows.append(ms)
        rows = np.asarray(rows)
        kern = {"split_ms": float(rows[:, 0].mean()), "convdet_ms": float(rows[:, 1].mean()),
                "detect_ms": float(rows[:, 2].mean())}
        for f in ("count", "anchor", "cls", "score", "box"):   # the profiled sequence is the production one
            step(nprof + 2)
            assert torch.equal(getattr(det_p, f), getattr(det, f)), f
        # the GEMM kernel on its own: back-to-back launches on pre-split planes (SQD_LAYOUT_SPLIT_NHWC input = the GEMM
        # launch only), one plane set per rotating feature set (3 x 115 MB > L2), CUDA events around the loop.  This is the
        # kernel's launch duration without the event gaps of the profiling twin (which serialise the chain and expose
        # the launch ramp that programmatic dependent launch hides in the real step).
        import ctypes as C
        gh, gw = shp.grid_hw
        cout = shp.out_channels
        pbytes = lib.sqd_convdet_split_bytes(B, shp.in_channels, gh, gw)
        planes = [torch.empty(pbytes, dtype=torch.uint8, device=dev) for _ in range(R)]
        st = _lib.stream_ptr(dev)
        for pl_, f in zip(planes, feats):
            _lib.check(lib.sqd_convdet_split_features(C.c_void_p(f.data_ptr()), _lib.LAYOUT_NCHW, B, shp.in_channels, gh, gw,
                                                      _lib.ptr(pl_), st), "sqd_convdet_split_features")
        ws_g = torch.empty(lib.sqd_convdet_workspace_bytes(B, shp.in_channels, gh, gw, cout, _lib.LAYOUT_SPLIT_NHWC,
                                                           _lib.CONV_TCGEN05_F16X3), dtype=torch.uint8, device=dev)
        pred_g = torch.empty((B, gh * gw * shp.anchors_per_grid, shp.num_classes + 5), device=dev)

        def gemm_only(i):
            _lib.check(lib.sqd_convdet_forward(_lib.ptr(planes[i % R]), _lib.LAYOUT_SPLIT_NHWC, _lib.ptr(packed), None,
                                               _lib.ptr(bias), B, shp.in_channels, gh, gw, cout, _lib.ptr(pred_g), _lib.ptr(ws_g),
                                               ws_g.numel(), _lib.CONV_TCGEN05_F16X3, st), "sqd_convdet_forward")
        for i in range(5):
            gemm_only(i)
        ng = max(20, min(K, 100))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(ng):
            gemm_only(i)
        g1.record()
        torch.cuda.synchronize()
        kern["convdet_alone_ms"] = g0.elapsed_time(g1) / ng
        del planes, ws_g, pred_g
        # decode + NMS on their own, at the sharded config's per-GPU size (BASELINE configs[2]: 2048 images over 2 GPUs
        # = 1024 per GPU) and on SURVEY 8d's pred-level synthetic set (background conf logit N(-3,1) + planted object
        # clusters, so NMS really suppresses): the HBM-bound shape of the path.  16 distinct images tiled on the device.
        Bd = args.decode_batch
        a64 = synth.anchor_table(shp)
        base = torch.from_numpy(synth.clustered_pred(shp, 16, 777, anchors=a64)).to(dev)
        pred_big = base.repeat((Bd + 15) // 16, 1, 1)[:Bd].contiguous()
        det_big = ops._alloc_detections(Bd, shp.top_k, dev)
        nrep = 12
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(nrep)]
        dense = None
        for i in range(nrep + 3):
            ev = evs[max(0, i - 3)]
            ev[0].record()
            ops.detect_from_pred(pred_big, anchors, shp.input_hw, shp.num_classes, shp.top_k, shp.nms_thresh, shp.score_thresh,
                                 two_phase=True, out=det_big)
            ev[1].record()
            dense = ops.decode_scores(pred_big, anchors, shp.input_hw, shp.num_classes, out=dense)
            ev[2].record()
        torch.cuda.synchronize()
        kern["detect_from_pred_big_ms"] = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
        kern["decode_scores_big_ms"] = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
        assert int(det_big.count.min()) > 0
        del pred_big, dense, det_big


    # ---- the other BASELINE.json configurations, driver-run at every rank count (max over ranks, whole-job values) -----
    def timed_steps(fn, n, warm):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b_.record()
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(a.elapsed_time(b_)) / n

    def use_graph_ok(g):
        return g["ms_per_step"] is not None

    extra = {}
    if not args.skip_extra:
        # configs[2]: 2048 KITTI images sharded by image over the ranks (strong scaling: total work fixed), no collective
        lo, hi = sdist.shard_range(args.config3_images, rank, world)
        B3 = hi - lo
        gen3 = torch.Generator(device=dev).manual_seed(4321 + rank)
        feat3 = torch.empty((B3, shp.in_channels, *shp.grid_hw), device=dev)
        for s0 in range(0, B3, 64):
            feat3[s0:s0 + 64] = torch.relu(torch.randn((min(64, B3 - s0), shp.in_channels, *shp.grid_hw), generator=gen3, device=dev))
        det3 = ops._alloc_detections(B3, shp.top_k, dev)
        ms3 = timed_steps(lambda i: ops.head_detect(feat3, weight, bias, anchors, shp.anchors_per_grid, shp.num_classes,
                                                    shp.input_hw, shp.top_k, shp.nms_thresh, shp.score_thresh, packed=packed,
                                                    out=det3), 3, 2)
        det3.check_status()
        assert int(det3.count.min()) >= 0
        extra["config3"] = {"workload": "BASELINE configs[2]: %d KITTI 1248x384 images sharded by image over %d GPU(s), "
                                        "features resident (%.1f GB per GPU > L2), no collective" % (args.config3_images, world, B3 * FEAT_BYTES_PER_IMAGE / 1e9),
                            "images_per_s": args.config3_images / (ms3 * 1e-3), "per_gpu": args.config3_images / (ms3 * 1e-3) / world,
                            "ms_per_step": ms3, "images_per_gpu": B3, "scaling": "strong", "steps": 3, "warmup": 2}
        del feat3, det3
        ops.workspace().clear()
        torch.cuda.empty_cache()

        # configs[4]: the high-density stress shape, 512 images over 8 GPUs = 64 per GPU (weak: 64 per rank at any N)
        sshp = synth.STRESS
        B5 = args.stress_batch
        gen5 = torch.Generator(device=dev).manual_seed(99 + rank)
        feat5 = [torch.relu(torch.randn((B5, sshp.in_channels, *sshp.grid_hw), generator=gen5, device=dev)) for _ in range(2)]
        w5_np, b5_np = synth.convdet_params(sshp, 4321)
        w5, b5 = torch.from_numpy(w5_np).to(dev), torch.from_numpy(b5_np).to(dev)
        packed5 = ops.pack_convdet_weights(w5)
        anchors5 = torch.from_numpy(synth.anchor_table(sshp).astype(np.float32)).to(dev)
        det5 = ops._alloc_detections(B5, sshp.top_k, dev)
        ms5 = timed_steps(lambda i: ops.head_detect(feat5[i % 2], w5, b5, anchors5, sshp.anchors_per_grid, sshp.num_classes,
                                                    sshp.input_hw, sshp.top_k, sshp.nms_thresh, sshp.score_thresh,
                                                    packed=packed5, out=det5), 6, 3)
        det5.check_status()
        flop5 = 2 * sshp.grid_hw[0] * sshp.grid_hw[1] * sshp.out_channels * 9 * sshp.in_channels
        extra["config5_stress"] = {"workload": "BASELINE configs[4]: 2496x768 input (67,392 anchors), 8 classes, top-256 per-class NMS, "
                                               "%d images per GPU (512 over 8 GPUs)" % B5,
                                   "images_per_s": world * B5 / (ms5 * 1e-3), "per_gpu": B5 / (ms5 * 1e-3), "ms_per_step": ms5,
                                   "images_per_gpu": B5, "scaling": "weak", "steps": 6, "warmup": 3,
                                   "tflops_algorithmic_per_gpu": B5 * flop5 / (ms5 * 1e-3) / 1e12,
                                   "kept_per_image_mean": float(det5.count.float().mean())}
        del feat5, det5, packed5
        ops.workspace().clear()
        torch.cuda.empty_cache()

        # configs[0]: the demo (demo.py:17-52), batch 1 per image: uint8 image on the HOST -> H2D -> sqd_preprocess (whiten
        # + resize) -> the stock backbone (cuDNN fp32, TF32 off; out of scope, timed because the demo runs it) -> the path
        # (sqd_head_detect_fused + sqd_boxes_postprocess behind Detector.detect) -> the reference's result dict on the host.
        if rank == 0:
            extra["config1_demo"] = demo_config(args, dev, ops, synth)
            ops.workspace().clear()
            torch.cuda.empty_cache()

        # configs[3]: the training step of the path, batch 20 per GPU: anchor matching + dense targets (a10-a13), ConvDet
        # forward (a1), loss forward + analytic backward (a14-a16), native wgrad / bias grad / dgrad (8f.2) and the NCCL
        # all-reduce of the flat gradient bucket (8e row 2; the full model's 2,082,120 floats, the head's segment launched
        # from the ConvDet backward before the dgrad GEMM is enqueued).  The backbone's own forward / backward is stock
        # PyTorch and out of scope: the step starts at the Fire11 features (which require grad, so the dgrad GEMM runs).
        from squeezedet_pytorch_b200 import config as sconfig, model as smodel, targets as stargets
        cfg = sconfig.make_config(shp, device=str(dev), dropout_prob=0.0)
        net = smodel.SqueezeDetWithLoss(cfg).to(dev)
        with torch.no_grad():
            net.base.convdet.weight.copy_(weight)
            net.base.convdet.bias.copy_(bias)
        net.train()

        class HeadWithLoss(torch.nn.Module):      # SqueezeDetWithLoss.forward (squeezedet.py:184-187) from the features on
            def __init__(self, full):
                super().__init__()
                self.base, self.loss = full.base, full.loss

            def forward(self, batch):
                pred = self.base.head(batch["features"])
                gt = batch["gt"]
                if isinstance(gt, tuple):      # (dense targets built on a side stream, the event that marks them ready)
                    gt, ready = gt
                    torch.cuda.current_stream(dev).wait_event(ready)
                return self.loss(pred, gt)

        class _FeatureGradSink(torch.autograd.Function):
            """Stands where the backbone's backward would: takes the feature gradient the dgrad GEMM wrote (channels_last
            memory) and ends the graph there.  Without it `tfeat` is a leaf and autograd's AccumulateGrad re-lays the
            115 MB gradient out as NCHW -- a 137 us strided copy per step (tools/train_profile.py) that belongs to neither
            the path nor a real training step, where the gradient goes on into the backbone."""
            @staticmethod
            def forward(ctx, x):
                return x.view_as(x)

            @staticmethod
            def backward(ctx, g):
                return None

        head = HeadWithLoss(net)
        bucket = sdist.bucket_for(net)
        matcher = stargets.AnchorMatcher(cfg.anchors, shp.num_classes, device=dev)
        cls_l, box_l = zip(*[synth.gt_boxes(shp, 100 + 1000 * rank + i) for i in range(B)])
        gt_packed = matcher.pack(list(box_l), list(cls_l))
        tfeat = feats[0].detach().clone().requires_grad_(True)

        # The reference builds the targets in DataLoader workers, i.e. beside the GPU work (datasets/base.py:52-76); here
        # the matcher (80 us, a latency chain on 160 CTAs) runs on a side stream beside the ConvDet forward and the loss
        # waits for its event -- inside the CUDA graph that is a fork / join.
        side = torch.cuda.Stream(device=dev)

        def targets_on_side_stream():
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gt = matcher.dense_targets(*gt_packed)
                ready = torch.cuda.Event()
                ready.record(side)
            gt.record_stream(cur)
            return gt, ready

        def train_iter(_i):
            tfeat.grad = None
            return sdist.train_step(head, {"features": _FeatureGradSink.apply(tfeat), "gt": targets_on_side_stream()}, bucket)

        nt = max(10, min(K, 50))
        ms_train_eager = timed_steps(train_iter, nt, 5)
        loss_val = float(train_iter(0)[0])
        assert np.isfinite(loss_val), loss_val
        real_distributed = bucket._distributed
        bucket._distributed = lambda: False           # the same step without the collective: the difference is what the
        ms_train_eager_local = timed_steps(train_iter, nt, 3)   # all-reduce costs the step (its exposed time)
        bucket._distributed = real_distributed
        # the step as ONE CUDA graph launch (dist.GraphedStep): the eager step is host bound, the replay is not
        graphed = {"ms_per_step": None, "ms_per_step_without_allreduce": None, "error": None}
        if not args.no_train_graph:
            try:
                g_full = sdist.GraphedStep(lambda: train_iter(0))
                graphed["ms_per_step"] = timed_steps(lambda _i: g_full(), nt, 3)
                g_loss = float(g_full()[0])
                assert abs(g_loss - loss_val) <= 1e-4 * abs(loss_val), (g_loss, loss_val)
                if world > 1:
                    bucket._distributed = lambda: False
                    g_local = sdist.GraphedStep(lambda: train_iter(0))
                    bucket._distributed = real_distributed
                    graphed["ms_per_step_without_allreduce"] = timed_steps(lambda _i: g_local(), nt, 3)
                    del g_local
                else:
                    graphed["ms_per_step_without_allreduce"] = graphed["ms_per_step"]
                del g_full
            except Exception as e:      # noqa: BLE001 -- report, keep the eager numbers
                bucket._distributed = real_distributed
                graphed["error"] = repr(e)[:300]
                torch.cuda.synchronize()
        ar_ms = None
        head_only = None
        if world > 1:
            def ar_only(_i):
                dist.all_reduce(bucket._buf, op=dist.ReduceOp.SUM)
            ar_ms = timed_steps(ar_only, 20, 5)
            # the same step with a bucket of the ConvDet head alone (497,736 floats): everything this path
            # produces is reduced beside the dgrad GEMM; what stays exposed in the full bucket is the backbone's segment,
            # whose gradients come from stock PyTorch after this path's backward
            if use_graph_ok(graphed):
                try:
                    for p_ in net.parameters():
                        p_.grad = None
                    hb = sdist.GradBucket(list(net.base.convdet.parameters()), early=list(net.base.convdet.parameters()))
                    net.base.grad_sink = hb

                    def head_iter(_i):
                        tfeat.grad = None
                        return sdist.train_step(head, {"features": _FeatureGradSink.apply(tfeat), "gt": targets_on_side_stream()}, hb)
                    gh_full = sdist.GraphedStep(lambda: head_iter(0))
                    t_with = timed_steps(lambda _i: gh_full(), nt, 3)
                    hb._distributed = lambda: False
                    gh_local = sdist.GraphedStep(lambda: head_iter(0))
                    t_without = timed_steps(lambda _i: gh_local(), nt, 3)
                    head_only = {"bucket_floats": int(hb.flat.numel()), "ms_per_step": t_with,
                                 "ms_per_step_without_allreduce": t_without, "exposed_allreduce_ms": max(0.0, t_with - t_without)}
                    del gh_full, gh_local, hb
                except Exception as e:      # noqa: BLE001
                    head_only = {"error": repr(e)[:300]}
                    torch.cuda.synchronize()
        use_graph = graphed["ms_per_step"] is not None
        ms_train = graphed["ms_per_step"] if use_graph else ms_train_eager
        ms_train_local = graphed["ms_per_step_without_allreduce"] if use_graph else ms_train_eager_local
        extra["train_step"] = {"workload": "BASELINE configs[3]: training step of the path at KITTI batch %d per GPU: matcher + targets (side "
                                           "stream, beside the forward), ConvDet forward, loss fwd+bwd, native wgrad/bias/dgrad (feature "
                                           "gradient handed over as channels_last memory), all-reduce of the flat gradient "
                                           "bucket (%d floats, head segment %d launched before the dgrad GEMM)" % (B, bucket.flat.numel(), bucket.early_numel),
                               "ms_per_step": ms_train, "images_per_s": world * B / (ms_train * 1e-3),
                               "mode": "one CUDA graph replay per step (dist.GraphedStep)" if use_graph else "eager",
                               "ms_per_step_without_allreduce": ms_train_local,
                               "exposed_allreduce_ms": max(0.0, ms_train - ms_train_local) if world > 1 else 0.0,
                               "exposed_allreduce_ms_head_bucket": (head_only or {}).get("exposed_allreduce_ms") if world > 1 else 0.0,
                               "exposed_note": "full-model bucket: the backbone's 1.58 M floats are reduced after this path's backward and "
                                               "have nothing to overlap with in this bench (the backbone's own backward is out of scope); "
                                               "the head's 497,736 floats -- every gradient this path produces -- are reduced beside the dgrad GEMM",
                               "eager": {"ms_per_step": ms_train_eager, "ms_per_step_without_allreduce": ms_train_eager_local,
                                         "note": "host bound: Python + autograd dispatch of ~40 launches"},
                               "graph_error": graphed["error"],
                               "head_only_bucket": head_only,
                               "allreduce_alone_ms": ar_ms, "nccl_ranks": world if world > 1 else 0,
                               "bucket_bytes": int(bucket._buf.numel() * 4), "loss": loss_val, "scaling": "weak"}
        del tfeat, net, head, bucket

    # ---- end to end with host buffers -------------------------------------------------------------------------
    host_feats = [torch.empty((B, shp.in_channels, *shp.grid_hw), dtype=torch.float32).pin_memory() for _ in range(2)]
    for hf, f in zip(host_feats, feats):
        hf.copy_(f.contiguous() if args.layout != "channels_last" else f.permute(0, 1, 2, 3).contiguous())
    host_dets = [ops.HostDetections(B, shp.top_k) for _ in range(2)]
    host_det = host_dets[0]
    h2d = host_feats[0].numel() * 4
    d2h = host_det.block.numel()      # ONE device-to-host copy of the whole result block per step

    e2e_seen = [0]

    def e2e_issue(i):
        # ONE public call per step: host features in, host detections out; H2D / kernels / D2H pipelined per image group
        # inside.  The serving form (sync=False, two slots): step i+1 is issued before step i's result is read, so the
        # PCIe link stays busy; EVERY step's result is read on the host (e2e_read) before its slot is reused.
        return ops.head_detect_host(host_feats[i % 2], weight, bias, anchors, shp.anchors_per_grid, shp.num_classes,
                                    shp.input_hw, shp.top_k, shp.nms_thresh, shp.score_thresh, packed=packed,
                                    out=host_dets[i % 2], chunk_images=args.e2e_chunk, sync=False, slot=i % 2)

    def e2e_read(det):
        det.wait()
        e2e_seen[0] += int(det.count.sum())      # the caller consumes the step's result

    def e2e_run(n):
        pending = None
        for i in range(n):
            det = e2e_issue(i)
            if pending is not None:
                e2e_read(pending)
            pending = det
        if pending is not None:
            e2e_read(pending)

    Ke = max(3, min(K, 50))
    if args.skip_e2e:      # profiling passes only: keeps the ncu launch list free of the chunked e2e launches
        Ke = 1
    e2e_run(0 if args.skip_e2e else 3)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_run(Ke)
    t_e2e = time.perf_counter() - t0
    assert e2e_seen[0] > 0
    barrier()
    t_e2e = max_over_ranks(t_e2e * 1e3) * 1e-3
    e2e_value = world * B * Ke / t_e2e
    # the floor under e2e: the SAME pinned buffers streamed to the device by bare cudaMemcpyAsync calls, all ranks at once,
    # no kernels, no result copies -- what the host / PCIe side of this box can deliver to N GPUs concurrently
    copy_floor = None
    if not args.skip_e2e:
        dst = torch.empty_like(host_feats[0], device=dev)
        cs = torch.cuda.Stream(device=dev)
        for i in range(2):
            with torch.cuda.stream(cs):
                dst.copy_(host_feats[i % 2], non_blocking=True)
        cs.synchronize()
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(cs):
            for i in range(Ke):
                dst.copy_(host_feats[i % 2], non_blocking=True)
        cs.synchronize()
        t_copy = time.perf_counter() - t0
        barrier()
        t_copy = max_over_ranks(t_copy * 1e3) * 1e-3
        copy_floor = {"images_per_s": world * B * Ke / t_copy, "gb_per_s": world * h2d * Ke / t_copy / 1e9,
                      "note": "bare pinned H2D copies of the step's features on all %d rank(s) at once (no kernels): the ceiling "
                              "of any host-buffer path on this box" % world}
        del dst
    if sampler.ok:
        sampler.stop()

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) -----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        info = run_cpu_arm(args, budget_s=15.0, warmup=2, steps=1000)
        cpu = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}

    # ---- the reference's OWN GPU path on this B200 (rank 0, N=1): cuDNN conv (TF32 off = the reference's fp32) + ATen
    #      decode kernels + the per-image filter loop with torchvision nms, restated op for op in oracle/torch_port.py ----
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import torch_port
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        a_cpu = torch.from_numpy(synth.anchor_table(shp).astype(np.float32))[None]
        run_ref = lambda f: torch_port.detect(f, weight, bias, a_cpu, shp.num_classes, shp.input_hw, shp.top_k,  # noqa: E731
                                              shp.nms_thresh, shp.score_thresh)
        gpu_kind = ("port of the reference's PyTorch CUDA path (oracle/torch_port.py): cuDNN fp32 conv + ATen decode + "
                    "per-image argsort / torchvision.ops.nms loop with its host syncs")
        try:    # the reference's OWN classes on the GPU when oracle/_ref travelled: cfg.device = cuda, nothing else changed
            ref_gpu = make_reference_runner(shp, w_np, b_np, device=str(dev))
        except Exception as e:      # noqa: BLE001 -- e.g. torchvision without CUDA ops: keep the port
            ref_gpu = None
            print("reference classes on the GPU unavailable: %r" % (e,), file=sys.stderr)
        if ref_gpu is not None:
            with torch.no_grad():
                first = ref_gpu(feats[0])
            assert len(first) == B
            run_ref = ref_gpu
            gpu_kind = ("reference: the reference's own SqueezeDet (Identity backbone) + Detector.detect from oracle/_ref "
                        "with cfg.device = cuda (cuDNN fp32 conv, ATen decode kernels, per-image argsort / torchvision nms "
                        "loop with its host syncs and per-image .cpu() copies)")
        for i in range(2):
            run_ref(feats[i % R])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nref = 0
        while nref < 5 or time.perf_counter() - t0 < 3.0:
            run_ref(feats[nref % R])
            nref += 1
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / nref
        gpu_ref = {"value": B / dt, "unit": "images/s", "kind": gpu_kind,
                   "sample": "%d passes over one batch of %d resident feature maps, wall clock" % (nref, B),
                   "ms_per_step": dt * 1e3}

    if rank == 0:
        conv_s = kern["convdet_alone_ms"] * 1e-3
        achieved = B * FLOP_PER_IMAGE / conv_s / 1e12
        achieved_in_step = B * FLOP_PER_IMAGE / (kern["convdet_ms"] * 1e-3) / 1e12
        det_s = kern["detect_from_pred_big_ms"] * 1e-3
        dec_s = kern["decode_scores_big_ms"] * 1e-3
        Bd = args.decode_batch
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "KITTI 1248x384 eval-shape inference, batch %d per GPU (BASELINE configs[1]): "
                                   "78x24 grid, 9 anchors, 3 classes, top-64, NMS 0.4" % B,
                       "input": "Fire11 feature maps (B,768,24,78) fp32 %s, resident in HBM" % args.layout,
                       "l2": "3 rotating input sets (345 MB) + 115 MB of fp16 planes per step > 126 MB L2; no flush needed",
                       "parallelism": "image-sharded, %d process(es), no collective" % world,
                       "serving_loop": "%d slots (stream + workspace + output block each), steps alternate; every step "
                                       "runs pre-pass, GEMM, scan and tail in full" % S,
                       "arithmetic": "fp32 results; the ConvDet products are fp16x3 (two-term fp16 split of power-of-two "
                                     "scaled operands on tcgen05, fp32 accumulate), fp32-grade: rms 9e-7 vs float64"},
            "per_gpu": value / world,
            "single_stream": {"images_per_s": world * B * K / (ms_single * 1e-3), "ms_per_step": ms_single / K,
                              "note": "the same K steps issued on ONE stream (each step's chain strictly after the previous "
                                      "step's tail kernel); `value` alternates the steps over %d slots (own stream, "
                                      "workspace and output block each)" % S},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "bare_h2d_copy_floor": copy_floor,
                    "frac_of_copy_floor": (e2e_value / copy_floor["images_per_s"]) if copy_floor else None, "note": "sqd_head_detect_host: pinned host features -> H2D in groups of %d images overlapped with the kernels "
                            "-> D2H of the detections; serving loop with two slots: step i+1 is issued before step i's "
                            "result is waited for and read on the host, every step's result is read" % args.e2e_chunk},
            "gpu_launches": 4 * K,
            "kernels_per_step": ["split_nchw_cluster_kernel (max|x| + fp16 split, one pass)", "convdet_f16_pair_kernel<80,0,false,36>",
                                 "score_candidates_kernel<3>", "detect_from_candidates_kernel"],
            "kernel_ms": kern,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": ncu_traffic(NCU_STEP_PROFILE, "convdet_f16_pair_kernel") if B == 20 else None,
                         "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                         "sustained_note": "launch_ms comes from %d back-to-back launches (a sustained tensor load: the SM clock "
                                           "sits at ~1.6 GHz under sw_power_cap); `frac` is nevertheless taken against the BURST "
                                           "peak, frac_of_sustained_peak against the sustained one (%.1f TFLOP/s)"
                                           % (max(20, min(K, 100)), peaks["tflops_sustained"]),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, parsed at run time from "
                                           "the committed ncu --set full summary %s (B = 20 only; ncu flushes L2 before the "
                                           "launch, so this is the cold-cache figure: algorithmic bytes are 115.0 MB of planes "
                                           "+ 2.0 MB weights + 10.8 MB pred)" % NCU_STEP_PROFILE,
                         "kernel": "convdet_f16_pair_kernel<80,0,false,36>",
                         "peak_source": peaks["source"] + ", dense bf16 burst",
                         "launch_ms": kern["convdet_alone_ms"],
                         "launch_ms_source": "CUDA events around back-to-back launches of the GEMM kernel on pre-split planes "
                                             "(3 rotating sets > L2); ncu launch list: profiles/r02_launches_b20.txt",
                         "in_step": {"launch_ms": kern["convdet_ms"], "achieved": achieved_in_step,
                                     "frac": achieved_in_step / peaks["tflops"],
                                     "note": "between the stage events of sqd_head_detect_profile: the events serialise the "
                                             "chain, so this includes the launch ramp that the real step overlaps"},
                         "note": "algorithmic FLOPs 2*1872*72*6912 per image; the kernel issues 3 fp16 passes (MMAs N = 144 + 80 "
                                 "for the 72 channels, 116.2 tensor cycles per K step measured: profiles/r01_umma_rate_2cta.txt) for "
                                 "fp32-level accuracy, so frac <= 0.31 by construction"},
            "roofline_decode_nms": {"bound": "hbm", "achieved": Bd * PRED_BYTES_PER_IMAGE / det_s / 1e9,
                                    "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": Bd * PRED_BYTES_PER_IMAGE / det_s / 1e9 / peaks["hbm_gbs"],
                                    "traffic": ncu_traffic(NCU_DECODE_PROFILE, "score_candidates_kernel", "detect_from_candidates_kernel") if Bd == 1024 else None,
                                    "traffic_source": NCU_DECODE_PROFILE + " (cold-cache ncu run: scan 93.2 us at 5.9 TB/s = 0.91 of peak, 128-thread tail 21.6 us)",
                                    "kernel": "score_candidates_kernel<3> + detect_from_candidates_kernel",
                                    "images_per_s": Bd / det_s,
                                    "note": "sqd_detect_from_pred (scan + per-image tail: the two kernels the fused step runs "
                                            "after the GEMM) on %d images of SURVEY 8d's clustered pred set resident in HBM "
                                            "(BASELINE configs[2] per-GPU size); algorithmic bytes 539,136 per image" % Bd},
            "roofline_decode": {"bound": "hbm", "achieved": Bd * (PRED_BYTES_PER_IMAGE + DENSE_BYTES_PER_IMAGE) / dec_s / 1e9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": Bd * (PRED_BYTES_PER_IMAGE + DENSE_BYTES_PER_IMAGE) / dec_s / 1e9 / peaks["hbm_gbs"],
                                "traffic": ncu_traffic(NCU_DECODE_PROFILE, "decode_kernel") if Bd == 1024 else None,
                                "traffic_source": NCU_DECODE_PROFILE, "kernel": "decode_kernel<3>",
                                "note": "sqd_decode_scores: the dense SqueezeDet.forward contract (class_ids i64, scores, "
                                        "boxes) for %d images: 539,136 B read + 471,744 B written per image" % Bd},
            "clocks": clocks,
        }
        line.update(extra)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gpu_ref is not None:
            line["gpu_reference_baseline"] = gpu_ref
        emit_result(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--slots", type=int, default=2, help="slots (stream + workspace + output block) the device-resident serving loop alternates over")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "channels_last"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="CPU arm: time the oracle port even when oracle/_ref is present")
    ap.add_argument("--decode-batch", type=int, default=1024, help="images of the stand-alone decode/NMS roofline measurement")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: one e2e step only (the JSON line is then not a bench result)")
    ap.add_argument("--e2e-chunk", type=int, default=5, help="images per H2D/compute pipeline group of the e2e call")
    ap.add_argument("--skip-extra", action="store_true", help="skip the configs[2], [3], [4] measurements (profiling aid)")
    ap.add_argument("--no-train-graph", action="store_true", help="time the training step eagerly only")
    ap.add_argument("--config3-images", type=int, default=2048, help="BASELINE configs[2]: images sharded over the ranks")
    ap.add_argument("--stress-batch", type=int, default=64, help="BASELINE configs[4]: images per GPU (512 over 8 GPUs)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result.  Libraries print there too (NCCL's "NCCL version" banner
    # under NCCL_DEBUG=VERSION is a C printf), so file descriptor 1 is pointed at stderr for the whole run and the
    # result goes to a private duplicate of the original stdout.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
