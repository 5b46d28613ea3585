"""Drop-in mirror of src/engine/detector.py.  `Detector.detect` keeps the reference's call
contract (list of per-image dicts of numpy arrays with post-processed boxes) but does the whole
tail -- ConvDet, decode, top-k, per-class NMS, threshold, boxes_postprocess -- in CUDA kernels for
the whole batch and crosses to the host ONCE (the reference syncs >= 3C+3 times per image)."""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.utils.data

from . import ops


def _meta_record(image_meta, batch_size):
    """image_meta as collated by the DataLoader (dict of (B,...) tensors / lists) -> (B,10) float32
    records for sqd_boxes_postprocess.  Absent keys become identities.  boxes.py:138-168"""
    rec = np.zeros((batch_size, 10), dtype=np.float32)
    rec[:, 0:2] = 1.0
    g = lambda k: np.asarray(image_meta[k].cpu() if torch.is_tensor(image_meta[k]) else image_meta[k])  # noqa: E731
    if "scales" in image_meta:
        rec[:, 0:2] = g("scales").reshape(batch_size, 2)
    if "padding" in image_meta:
        p = g("padding").reshape(batch_size, 4)
        rec[:, 2], rec[:, 3] = p[:, 0], p[:, 2]
    if "crops" in image_meta:
        c = g("crops").reshape(batch_size, 4)
        rec[:, 4], rec[:, 5] = c[:, 0], c[:, 2]
    if "flipped" in image_meta:
        flipped = g("flipped").reshape(batch_size).astype(bool)
        size = g("drifted_size") if "drifted_size" in image_meta else g("orig_size")
        rec[:, 6] = np.where(flipped, size.reshape(batch_size, -1)[:, 1], 0)
    if "drifts" in image_meta:
        d = g("drifts").reshape(batch_size, 2)
        rec[:, 7], rec[:, 8] = d[:, 0], d[:, 1]
    return rec


class Detector(object):
    def __init__(self, model, cfg):
        self.model = model.to(cfg.device)
        self.model.eval()
        self.cfg = cfg
        # cfg.cuda_graph = True: backbone + path of a batch shape are captured once into a CUDA graph and replayed --
        # the batch-1 demo (demo.py:17-52) is launch bound in the stock backbone (~60 small kernels, 1.8 ms eager)
        self.use_graph = bool(getattr(cfg, "cuda_graph", False))
        self._graphs = {}

    # -- the fast path: features -> final detections, one host crossing ------------------------------
    @torch.no_grad()
    def detect_batch(self, batch, postprocess=True):
        """Device-side result for a batch {'image': (B,3,H,W)} (+ optional 'image_meta'):
        (Detections with the boxes still in network-input coordinates, (B,10) postprocess records or None)."""
        image = batch["image"]
        det = self._detect_graphed(image) if self.use_graph and image.is_cuda else self._detect_eager(image)
        meta = None
        if postprocess and "image_meta" in batch and batch["image_meta"]:
            meta = torch.from_numpy(_meta_record(batch["image_meta"], det.count.shape[0])).to(image.device)
        return det, meta

    def _detect_eager(self, image, out=None):
        cfg, base = self.cfg, self.model.base
        feat = base.features(image)
        anchors = self.model.resolver._anchors_on(feat.device)
        return ops.head_detect(feat, base.convdet.weight, base.convdet.bias, anchors, cfg.anchors_per_grid,
                               cfg.num_classes, cfg.input_size, cfg.keep_top_k, cfg.nms_thresh, cfg.score_thresh,
                               packed=base.packed_weights(), algo=base.conv_algo, out=out)

    def _detect_graphed(self, image):
        """One CUDA graph per (input shape, weight version): replay = copy the batch into the graph's input buffer + one
        launch.  The returned Detections are the graph's own output block: valid until the next call with this shape."""
        w = self.model.base.convdet.weight
        key = (tuple(image.shape), image.dtype, str(image.device), w._version, w.data_ptr())
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 8:            # stale shapes / weight versions
                self._graphs.clear()
            static_in = torch.empty_like(image)
            static_in.copy_(image)
            det = ops._alloc_detections(image.shape[0], self.cfg.keep_top_k, image.device)
            cur = torch.cuda.current_stream(image.device)
            side = torch.cuda.Stream(device=image.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(3):                # cuDNN algorithm selection, workspaces, packed weights: before capture
                    self._detect_eager(static_in, out=det)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._detect_eager(static_in, out=det)
            entry = self._graphs[key] = (graph, static_in, det)
        graph, static_in, det = entry
        static_in.copy_(image)
        graph.replay()
        return det

    @torch.no_grad()
    def detect_packed(self, batch):
        """(packed (B,k,6) [class, score, x1,y1,x2,y2] with boxes_postprocess applied, count (B,)) on the HOST:
        two device->host copies for the whole batch."""
        from . import results
        det, meta = self.detect_batch(batch)
        return results.to_host(det, meta)

    @torch.no_grad()
    def detect(self, batch):
        """Reference contract, detector.py:20-50: list of {'class_ids','scores','boxes','image_meta'}
        (or {'image_meta'} only when an image keeps nothing)."""
        from . import results
        packed, count = self.detect_packed(batch)
        meta_in = batch.get("image_meta", {}) or {}
        metas = [{k: (v[b].cpu().numpy() if torch.is_tensor(v) else v[b]) for k, v in meta_in.items()}
                 for b in range(count.shape[0])]
        return results.unpack(packed, count, metas)

    def detect_dataset(self, dataset):
        """detector.py:52-85: DataLoader loop with data / net timers (I/O glue, stock PyTorch)."""
        from .compat import DataWrapper
        start_time = time.time()
        loader = torch.utils.data.DataLoader(DataWrapper(dataset), batch_size=self.cfg.batch_size,
                                             num_workers=self.cfg.num_workers, pin_memory=True)
        results, end = [], time.time()
        data_t = net_t = 0.0
        for iter_id, batch in enumerate(loader):
            for k in batch:
                if "image_meta" not in k:
                    batch[k] = batch[k].to(device=self.cfg.device, non_blocking=True)
            data_t, end = time.time() - end, time.time()
            results.extend(self.detect(batch))
            net_t, end = time.time() - end, time.time()
            if iter_id % self.cfg.print_interval == 0:
                print("eval: [{0}/{1}] | data {2:.3f}s | net {3:.3f}s".format(iter_id, len(loader), data_t, net_t))
        tpi = (time.time() - start_time) / max(1, len(dataset))
        print("Elapsed {:.2f}min ({:.1f}ms/image, {:.1f}frames/s)".format(tpi * len(dataset) / 60., tpi * 1000., 1 / tpi))
        return results

    # -- the reference's per-image filter contract ----------------------------------------------------
    def filter(self, det):
        """detector.py:87-122 for ONE image: dict of (A,), (A,), (A,4) device tensors -> dict of kept
        (n,), (n,), (n,4) device tensors, or None.  (Also returns 'anchor_idx': the kept anchor ids.)"""
        cfg = self.cfg
        out = ops.topk_nms(det["class_ids"][None].to(torch.int64), det["scores"][None].float(),
                           det["boxes"][None].float(), cfg.num_classes, cfg.keep_top_k, cfg.nms_thresh,
                           cfg.score_thresh)
        n = int(out.count[0].item())
        if n == 0:
            return None
        return {"class_ids": out.cls[0, :n].to(torch.int64), "scores": out.score[0, :n], "boxes": out.box[0, :n],
                "anchor_idx": out.anchor[0, :n].to(torch.int64)}
