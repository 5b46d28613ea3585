"""Multi-GPU plumbing.  One process per GPU.

Inference shards by image: rank r owns images [r*B/n, (r+1)*B/n) and runs the whole path on its
slice; there is NO collective on the inference path.  Training is data parallel over images with
ONE collective per step: an all-reduce(sum) of a single flat fp32 gradient bucket, replacing the
reference's per-step parameter broadcast + reduce-to-GPU-0 (src/utils/data_parallel.py:93-101)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, world, local_rank); a no-op single process when the variables are absent."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)   # binds the communicator to this rank's GPU (no guessing)
            # NCCL writes its debug output (the "NCCL version" banner included, at any level >= VERSION) to stdout,
            # which carries bench.py's one JSON line: route it to stderr instead of silencing it.
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous, balanced image range of this rank: sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world):
    """Slice every batched tensor / list of a batch dict to this rank's images."""
    n = None
    for v in batch.values():
        if torch.is_tensor(v):
            n = v.shape[0]
            break
    lo, hi = shard_range(n, rank, world)
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v) or isinstance(v, list):
            out[k] = v[lo:hi]
        elif isinstance(v, dict):
            out[k] = {kk: vv[lo:hi] for kk, vv in v.items()}
        else:
            out[k] = v
    return out


class GradBucket:
    """All parameter gradients as views into one flat fp32 buffer -> one all-reduce per step.

    `early`: the parameters whose gradients autograd finishes FIRST -- the ConvDet head (its wgrad / bias-grad kernels
    run before anything of the backbone's backward).  They sit at the front of the buffer and the
    all-reduce of that segment is launched (async) from the hook of the last of them, so it travels over NVLink while
    the backbone's backward still runs; `allreduce_sum` / `allreduce_mean` then reduce the rest and wait for both.  The
    result is the same as one all-reduce of the whole buffer.

    Two leading slots ride along with the FIRST segment (no extra collective, no host sync): `loss_slot` (this rank's
    SUM of per-image losses -- known before backward starts) and `count_slot` (this rank's image count).  After the
    all-reduce they hold the global loss sum and the global batch size, and `allreduce_sum(normalize=True)` divides the
    gradients by that count on the device -- the reference's whole-batch `loss.mean()` (trainer.py:43) for ANY split
    of the batch, uneven or with empty shards.  Travelling with the early segment they are reduced beside the dgrad
    GEMM; a bucket that holds only the head then has no collective left after backward.  Every rank issues the same
    sequence of collectives whether or not its hooks fired."""

    def __init__(self, params, early=()):
        early = [p for p in early if p.requires_grad]
        seen = {id(p) for p in early}
        self.params = early + [p for p in params if p.requires_grad and id(p) not in seen]
        self.early_numel = sum(p.numel() for p in early)
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self._buf = torch.zeros(total + 4, dtype=torch.float32, device=dev)   # [loss, count, 0, 0 | gradients]: 16-byte aligned
        self.loss_slot = self._buf[0:1]
        self.count_slot = self._buf[1:2]
        self.flat = self._buf[4:]                       # the gradients
        self._early_seg = self._buf[:4 + self.early_numel]      # slots + head gradients: the first all-reduce
        self._rest_seg = self._buf[4 + self.early_numel:]       # everything else (may be empty)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._n_early, self._pending, self._early_work = len(early), 0, None
        self._early_ids = {id(p) for p in early}
        for p in early:
            p.register_post_accumulate_grad_hook(self._early_ready)

    @staticmethod
    def _distributed():
        return dist.is_initialized() and dist.get_world_size() > 1

    def _early_ready(self, _param):
        if self._pending <= 0:
            return                  # not armed (zero() was not called for this step) or already launched
        self._pending -= 1
        if self._pending == 0 and self._distributed():
            self._early_work = dist.all_reduce(self._early_seg, op=dist.ReduceOp.SUM, async_op=True)

    def take_early(self, grads):
        """Fast path used by the ConvDet backward (model._ConvDetFn): `grads` = {id(param): (param, grad)} for exactly the
        early parameters, handed over BEFORE the feature gradient is enqueued.  They are added into the bucket and the
        early all-reduce is launched right away, so on the GPU it runs beside the dgrad GEMM and the backbone backward
        instead of behind them.  Returns False (nothing taken) when the bucket is not armed or the set does not match;
        autograd then accumulates the gradients as usual and the hooks launch the collective."""
        if self._pending <= 0 or set(grads) != self._early_ids or any(g is None for _, g in grads.values()):
            return False
        for p, g in grads.values():
            p.grad.add_(g.view_as(p.grad))
        self._pending = 0
        if self._distributed():
            self._early_work = dist.all_reduce(self._early_seg, op=dist.ReduceOp.SUM, async_op=True)
        return True

    def zero(self):
        """Start of a step: clear the gradients (and the loss / count slots) and arm the early all-reduce."""
        self._buf.zero_()
        self._pending, self._early_work = self._n_early, None

    def allreduce_sum(self, normalize=False):
        """Sum the bucket over the ranks.  normalize: divide the gradients by the all-reduced `count_slot` (device-side;
        the caller put its image count there and back-propagated the SUM of its per-image losses)."""
        if self._distributed():
            armed = self._pending > 0 or self._early_work is not None
            if self._n_early and armed:
                if self._early_work is None:    # the hooks never fired (empty shard: no backward): same collectives anyway
                    self._early_work = dist.all_reduce(self._early_seg, op=dist.ReduceOp.SUM, async_op=True)
                work = None
                if self._rest_seg.numel() > 0:
                    work = dist.all_reduce(self._rest_seg, op=dist.ReduceOp.SUM, async_op=True)
                self._early_work.wait()
                if work is not None:
                    work.wait()
                self._early_work = None
            else:
                dist.all_reduce(self._buf, op=dist.ReduceOp.SUM)
        self._pending = 0
        if normalize:
            self.flat.div_(self.count_slot)

    def allreduce_mean(self, world=None):
        """sum over ranks then scale by 1/world: for callers whose ranks back-propagate the MEAN loss of EQUAL-sized
        shards.  (train_step uses the count-weighted form instead, which is exact for any split.)"""
        distributed = self._distributed()
        self.allreduce_sum(normalize=False)
        if distributed:
            self.flat.mul_(1.0 / (world or dist.get_world_size()))
        return None


def bucket_for(model):
    """GradBucket of a SqueezeDetWithLoss / SqueezeDetBase-holding module with the ConvDet head as the early segment."""
    base = model.base if hasattr(model, "base") else model
    bucket = GradBucket(model.parameters(), early=base.convdet.parameters())
    base.grad_sink = bucket
    return bucket


def _batch_images(batch):
    for v in batch.values():
        if torch.is_tensor(v):
            return int(v.shape[0])
    return 0


def train_step(model, batch, bucket, optimizer=None, grad_norm=None):
    """One data-parallel step on this rank's shard of the batch: the body of Trainer.run_epoch (src/engine/
    trainer.py:42-48) with the reference's per-step parameter broadcast + gradient reduce-to-GPU-0
    (src/utils/data_parallel.py:93-101) replaced by the bucket's all-reduce.  `optimizer.zero_grad()` is
    `bucket.zero()` here (the gradients are views into the bucket and must stay allocated).

    The reference averages the per-image losses over the WHOLE batch (trainer.py:43), so each rank back-propagates the
    SUM of its images' losses and the summed gradients are divided by the global image count, which travels in the
    bucket: shards of different sizes weigh every image equally, and a rank with an empty shard contributes zeros
    (it skips forward / backward but still takes part in the collectives).
    Returns (mean loss over the global batch as a 0-d device tensor, this shard's per-image statistics dict or {})."""
    bucket.zero()
    n_local = _batch_images(batch)
    bucket.count_slot.fill_(float(n_local))
    stats = {}
    if n_local > 0:
        loss, stats = model(batch)
        loss_sum = loss.sum()
        bucket.loss_slot.copy_(loss_sum.detach().reshape(1))    # before backward: the slots travel with the early segment
        loss_sum.backward()         # the head's all-reduce is launched from inside, as soon as its gradients exist
    bucket.allreduce_sum(normalize=True)
    if grad_norm:
        torch.nn.utils.clip_grad_norm_(bucket.params, grad_norm)
    if optimizer is not None:
        optimizer.step()
    return (bucket.loss_slot / bucket.count_slot).reshape(()), stats


class GraphedStep:
    """One training step as a CUDA graph: `fn()` (e.g. `lambda: train_step(model, batch, bucket)`) is run a few times on
    a side stream, captured once -- forward, native backward, the bucket's NCCL all-reduces and the normalisation are
    all stream work -- and replayed with ONE launch per step.  The eager step of the head path is host bound (about
    forty kernel launches, three of them collectives, behind Python and autograd dispatch); the replay runs at the
    speed of its kernels and the all-reduce of the head segment overlaps the dgrad GEMM as captured.

    Contract of CUDA graphs: every tensor `fn` reads must keep its address -- feed new data with `.copy_()` into the
    tensors `fn` closed over -- and `fn` must not synchronise with the host.  The value `fn` returned during capture
    (`.result`) is refreshed by every replay."""

    def __init__(self, fn, warmup=3):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result
