"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the host
logic (sharding, gradient bucket, config) works, and the product path refuses to run without CUDA."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "sqdet_b200.h")).read()
    return sorted(set(re.findall(r"SQD_API[^;]*?\b(sqd_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from squeezedet_pytorch_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sqdet_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes table covers the whole header
    assert lib.sqd_abi_version() == 4


def test_size_queries_need_no_gpu():
    from squeezedet_pytorch_b200 import _lib
    lib = _lib.load()
    # KITTI: Cout 72 -> padded 80 rows, two fp16 terms, two layouts (CTA-pair and 1-CTA kernels), 256-byte header
    assert lib.sqd_convdet_packed_weight_bytes(72, 768) == 256 + 2 * 2 * 80 * 9 * 768 * 2
    # workspace holds the two fp16 NHWC planes of the batch
    assert lib.sqd_convdet_workspace_bytes(20, 768, 24, 78, 72, 0, 0) >= 2 * 20 * 768 * 24 * 78 * 2
    assert lib.sqd_convdet_split_bytes(20, 768, 24, 78) >= 2 * 20 * 768 * 24 * 78 * 2
    assert lib.sqd_loss_workspace_bytes(20, 16848) > 0


def test_product_path_has_no_cpu_fallback():
    from squeezedet_pytorch_b200 import _lib, ops, synth
    shp = synth.TINY
    pred = torch.zeros((1, shp.num_anchors, 8))
    anchors = torch.zeros((shp.num_anchors, 4))
    with pytest.raises(_lib.SqdError):
        ops.decode_scores(pred, anchors, shp.input_hw, 3)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "squeezedet-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "from oracle" not in text and "import oracle" not in text, f


def test_synth_is_deterministic_and_shaped():
    from squeezedet_pytorch_b200 import synth
    a = synth.clustered_pred(synth.TINY, 2, 3)
    b = synth.clustered_pred(synth.TINY, 2, 3)
    assert np.array_equal(a, b) and a.shape == (2, 540, 8) and a.dtype == np.float32
    f = synth.features(synth.TINY, 1, 0)
    assert f.shape == (1, 768, 6, 10) and f.min() >= 0
    assert synth.KITTI.num_anchors == 16848 and synth.STRESS.num_anchors == 67392
    w, bias = synth.convdet_params(synth.KITTI, 1)
    assert w.shape == (72, 768, 3, 3) and bias.shape == (72,)


def test_config_and_state_dict_keys_match_reference_names():
    from squeezedet_pytorch_b200 import config, model
    cfg = config.kitti_config(device="cpu")
    for field in ("num_classes", "num_anchors", "anchors_per_grid", "anchors", "input_size", "arch", "dropout_prob",
                  "keep_top_k", "nms_thresh", "score_thresh", "device", "class_loss_weight",
                  "positive_score_loss_weight", "negative_score_loss_weight", "bbox_loss_weight"):
        assert hasattr(cfg, field)
    net = model.SqueezeDetWithLoss(cfg)
    keys = set(net.state_dict())
    assert {"base.convdet.weight", "base.convdet.bias", "base.features.0.weight", "base.features.3.squeeze.weight",
            "base.features.12.expand3x3.bias"} <= keys
    assert net.base.convdet.weight.shape == (72, 768, 3, 3)
    assert sum(p.numel() for p in net.parameters()) == 2082120
    plus = model.SqueezeDetBase(config.kitti_config(device="cpu", arch="squeezedetplus"))
    assert plus.convdet.weight.shape == (72, 512, 3, 3)


def test_shard_range_partitions_exactly():
    from squeezedet_pytorch_b200 import dist as sdist
    for total in (20, 2048, 7, 1):
        for world in (1, 2, 4, 8):
            spans = [sdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from squeezedet_pytorch_b200 import dist as sdist
rank, world, _ = sdist.init_from_env(backend="gloo")
assert world == 2
torch.manual_seed(0)
lin = torch.nn.Linear(8, 4)
bucket = sdist.GradBucket(lin.parameters())
batch = {"image": torch.arange(10 * 8, dtype=torch.float32).view(10, 8), "ids": list(range(10)),
         "image_meta": {"index": torch.arange(10)}}
mine = sdist.shard_batch(batch, rank, world)
assert mine["image"].shape[0] == 5 and mine["ids"] == list(range(rank * 5, rank * 5 + 5))
assert mine["image_meta"]["index"].tolist() == list(range(rank * 5, rank * 5 + 5))
bucket.zero()
lin(mine["image"]).mean().backward()
bucket.allreduce_mean()
# reference: the same model on the full batch in one process
ref = torch.nn.Linear(8, 4); ref.load_state_dict(lin.state_dict())
ref(batch["image"]).mean().backward()
flat_ref = torch.cat([p.grad.flatten() for p in ref.parameters()])
assert torch.allclose(bucket.flat, flat_ref, rtol=1e-5, atol=1e-6), (bucket.flat, flat_ref)
gathered = [torch.zeros_like(bucket.flat) for _ in range(world)]
dist.all_gather(gathered, bucket.flat)
assert torch.equal(gathered[0], gathered[1])
# early segment: the "head" parameters are reduced from a gradient hook while backward is still running
torch.manual_seed(1)
net = torch.nn.Sequential(torch.nn.Linear(8, 6), torch.nn.ReLU(), torch.nn.Linear(6, 4))
head = list(net[2].parameters())
b2 = sdist.GradBucket(net.parameters(), early=head)
assert b2.early_numel == 6 * 4 + 4 and b2.params[0] is head[0]
launched = []
orig = dist.all_reduce
def spy(t, *a, **k):
    launched.append(t.numel())
    return orig(t, *a, **k)
dist.all_reduce = spy
for step in range(2):
    b2.zero()
    net(mine["image"] / 80.0).mean().backward()
    assert launched[-1] == b2.early_numel + 4 and len(launched) == 3 * step + 1     # fired from the hook (+ loss / count slots)
    b2.allreduce_mean()
    assert launched[-1] == b2.flat.numel() - b2.early_numel and len(launched) == 3 * step + 2
    launched.append(0)
dist.all_reduce = orig
ref2 = torch.nn.Sequential(torch.nn.Linear(8, 6), torch.nn.ReLU(), torch.nn.Linear(6, 4)); ref2.load_state_dict(net.state_dict())
ref2(batch["image"] / 80.0).mean().backward()
want = torch.cat([p.grad.flatten() for p in list(ref2[2].parameters()) + list(ref2[0].parameters())])
assert torch.allclose(b2.flat, want, rtol=1e-5, atol=1e-7), (b2.flat, want)
# train_step: count-weighted all-reduce == the reference's whole-batch loss.mean() for UNEVEN and EMPTY shards
class PerImage(torch.nn.Module):            # stands in for SqueezeDetWithLoss: batch dict -> (per-image loss, stats)
    def __init__(self):
        super().__init__()
        self.base = torch.nn.Module()
        self.base.convdet = torch.nn.Linear(6, 4)
        self.body = torch.nn.Linear(8, 6)
    def forward(self, batch):
        l = (self.base.convdet(torch.relu(self.body(batch["image"]))) ** 2).sum(1)
        return l, {"loss": l}
for total in (7, 1, 10):
    torch.manual_seed(5)
    m = PerImage()
    bk = sdist.bucket_for(m)
    assert bk.early_numel == 6 * 4 + 4
    full = {"image": torch.randn(total, 8, generator=torch.Generator().manual_seed(total))}
    part = sdist.shard_batch(full, rank, world)
    if total == 1:
        assert part["image"].shape[0] == (1 if rank == 0 else 0)      # rank 1 owns nothing
    loss, _ = sdist.train_step(m, part, bk)
    m2 = PerImage(); m2.load_state_dict(m.state_dict())
    l2, _ = m2(full)
    l2.mean().backward()
    want = torch.cat([p.grad.flatten() for p in list(m2.base.convdet.parameters()) + list(m2.body.parameters())])
    assert torch.isfinite(bk.flat).all()
    assert torch.allclose(bk.flat, want, rtol=1e-5, atol=1e-6), (total, bk.flat, want)
    assert torch.allclose(loss, l2.mean().detach(), rtol=1e-5), (total, loss, l2.mean())
# take_early: a head whose backward hands its weight / bias gradients to the bucket BEFORE computing the input gradient
# (what model._ConvDetFn does): the early all-reduce is launched from inside backward, the result is unchanged
class HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, sink):
        ctx.save_for_backward(x, w, b); ctx.sink = sink
        return x @ w.t() + b
    @staticmethod
    def backward(ctx, g):
        x, w, b = ctx.saved_tensors
        gw, gb = g.t() @ x, g.sum(0)
        order.append("wgrad")
        if ctx.sink.take_early({id(w): (w, gw), id(b): (b, gb)}):
            gw = gb = None
        order.append("dgrad")
        return g @ w, gw, gb, None
class EarlyHead(PerImage):
    def forward(self, batch):
        l = (HeadFn.apply(torch.relu(self.body(batch["image"])), self.base.convdet.weight, self.base.convdet.bias,
                          self.base.grad_sink) ** 2).sum(1)
        return l, {"loss": l}
order = []
def spy2(t, *a, **k):
    order.append("allreduce%d" % t.numel())
    return orig(t, *a, **k)
dist.all_reduce = spy2
torch.manual_seed(9)
m = EarlyHead()
bk = sdist.bucket_for(m)
assert m.base.grad_sink is bk
full = {"image": torch.randn(9, 8, generator=torch.Generator().manual_seed(3))}
loss, _ = sdist.train_step(m, sdist.shard_batch(full, rank, world), bk)
assert order[:3] == ["wgrad", "allreduce%d" % (bk.early_numel + 4), "dgrad"], order      # launched between wgrad and dgrad
assert order.count("allreduce%d" % (bk.early_numel + 4)) == 1, order
dist.all_reduce = orig
m2 = PerImage(); m2.load_state_dict(m.state_dict())
l2, _ = m2(full)
l2.mean().backward()
want = torch.cat([p.grad.flatten() for p in list(m2.base.convdet.parameters()) + list(m2.body.parameters())])
assert torch.allclose(bk.flat, want, rtol=1e-5, atol=1e-6), (bk.flat, want)
# not armed (no bucket.zero()): autograd accumulates as usual
for p_ in m.parameters():
    p_.grad = None
lz, _ = m(full)
lz.mean().backward()
assert torch.allclose(m.base.convdet.weight.grad, m2.base.convdet.weight.grad, rtol=1e-5, atol=1e-6)
dist.destroy_process_group()
print("OK", rank)
'''


def test_gloo_world_size_2_shard_and_gradient_allreduce(tmp_path):
    """N>1 host path on CPU: image sharding + ONE flat-bucket all-reduce reproduce the single-process
    full-batch gradient (the data-parallel contract of SURVEY 8e)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = 29500 + (os.getpid() % 1000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o


def test_option_table_is_set_through_the_abi_not_the_environment(monkeypatch):
    """Developer options: seeded from the environment once, then only sqd_set_option changes them (no getenv per call)."""
    import ctypes as C
    from squeezedet_pytorch_b200 import _lib
    lib = _lib.load()
    v = C.c_int(-7)
    assert lib.sqd_get_option(b"SQD_SPLIT_TWO_PASS", C.byref(v)) == 0 and v.value == 0
    monkeypatch.setenv("SQD_SPLIT_TWO_PASS", "1")            # too late: the table was filled on first use
    assert lib.sqd_get_option(b"SQD_SPLIT_TWO_PASS", C.byref(v)) == 0 and v.value == 0
    with _lib.option("SQD_SPLIT_TWO_PASS", 1):
        assert lib.sqd_get_option(b"SQD_SPLIT_TWO_PASS", C.byref(v)) == 0 and v.value == 1
    assert lib.sqd_get_option(b"SQD_SPLIT_TWO_PASS", C.byref(v)) == 0 and v.value == 0
    assert lib.sqd_set_option(b"SQD_NO_SUCH_OPTION", 1) == -5 and lib.sqd_last_error().startswith(b"sqd_set_option")
    assert lib.sqd_set_option(None, 1) == -1 and lib.sqd_get_option(b"SQD_NO_PDL", None) == -1
    # the shipped library carries no trace hook and never reads the environment per call
    blob = open(_lib.LIB_PATH, "rb").read()
    assert b"SQD_F16_TRACE\x00" not in blob


def _abi_call(lib, name, ptr_value, int_value, size_value=0):
    """Call an int-returning entry point with every pointer = ptr_value, every int = int_value, doubles 0.5."""
    import ctypes as C
    from squeezedet_pytorch_b200 import _lib
    _, args = _lib.SIGNATURES[name]
    vals = []
    for a in args:
        if a is C.c_void_p:
            vals.append(ptr_value)
        elif hasattr(a, "_type_") and not isinstance(a._type_, str):       # POINTER(...)
            vals.append(C.cast(ptr_value, a) if ptr_value else None)
        elif a is C.c_double:
            vals.append(0.5)
        elif a is C.c_size_t:
            vals.append(size_value)
        else:
            vals.append(int_value)
    rc = getattr(lib, name)(*vals)
    return rc, lib.sqd_last_error().decode()


def test_abi_rejects_bad_arguments_before_touching_the_device():
    """Error behaviour of the boundary (int return codes, message through sqd_last_error(), never throw,
    never crash): every compute entry point refuses NULL pointers with SQD_E_NULL and nonsense shapes with a negative
    code, naming itself in the message -- decided on the host, so it runs here without a GPU.  Entry points that get
    past validation report the CUDA error (> 0) instead of computing on the CPU: there is no fallback."""
    import ctypes as C
    from squeezedet_pytorch_b200 import _lib
    lib = _lib.load()
    compute = [n for n, (res, args) in _lib.SIGNATURES.items()
               if res is C.c_int and args and n not in ("sqd_set_option", "sqd_get_option")]
    assert len(compute) >= 20
    for name in compute:
        rc, msg = _abi_call(lib, name, None, 1)
        assert rc == -1, (name, rc, msg)                                    # SQD_E_NULL
        assert msg.startswith(name), (name, msg)
        rc, msg = _abi_call(lib, name, 0x10000, -1)
        if name.endswith("_status"):                                        # one pointer, no shape: goes to the device
            assert rc > 0, (name, rc, msg)
            continue
        assert rc in (-2, -5), (name, rc, msg)                              # SQD_E_SHAPE / SQD_E_UNSUPPORTED
        assert msg.startswith(name), (name, msg)
    # specific contracts
    kitti = dict(batch=2, cin=768, gh=24, gw=78, cout=72)
    need = lib.sqd_head_detect_workspace_bytes(kitti["batch"], kitti["cin"], kitti["gh"], kitti["gw"], kitti["cout"], 0, _lib.CONV_TCGEN05_F16X3)
    assert need > 2 * 768 * 24 * 78 * 4                                     # at least the two fp16 planes of the features
    p = C.c_void_p(0x10000)
    rc = lib.sqd_head_detect_fused(p, 0, p, p, p, p, 2, 768, 24, 78, 9, 3, 384, 1248, 64, 0.4, 0.3, p, p, p, p, p, p, need - 1,
                                   _lib.CONV_TCGEN05_F16X3, None)
    assert rc == -3 and b"workspace too small" in lib.sqd_last_error()       # SQD_E_WORKSPACE
    rc = lib.sqd_head_detect_fused(p, 0, p, p, p, p, 2, 768, 24, 78, 9, 3, 384, 1248, 64, 0.4, 0.3, p, p, p, p, p,
                                   C.c_void_p(0x10004), need, _lib.CONV_TCGEN05_F16X3, None)
    assert rc in (-3, -4), rc                                               # misaligned workspace
    rc = lib.sqd_convdet_pack_weights(p, 72, 700, p, None)                  # Cin not a multiple of 64
    assert rc == -2 and b"multiple of 64" in lib.sqd_last_error()
    rc = lib.sqd_match_anchors(p, p, 1, 257, p, 100, p, p, None)            # more ground-truth boxes than the kernel holds
    assert rc == -2 and b"gmax" in lib.sqd_last_error()
    rc = lib.sqd_preprocess(p, 7, 1, 10, 10, C.cast(p, C.POINTER(C.c_float)), C.cast(p, C.POINTER(C.c_float)), 5, 5, p, None)
    assert rc == -5 and b"dtype" in lib.sqd_last_error()                    # SQD_E_UNSUPPORTED
    # an empty batch is legal everywhere and enqueues nothing (pointers may be NULL)
    for name in ("sqd_decode_scores", "sqd_topk_nms", "sqd_detect_from_pred", "sqd_boxes_postprocess", "sqd_pack_results",
                 "sqd_preprocess"):
        _, args = _lib.SIGNATURES[name]
        rc, msg = _abi_call(lib, name, None, 0)
        assert rc == 0, (name, rc, msg)
