"""Run under torchrun with N >= 2 GPUs: a data-parallel training step (dist.train_step: sharded batch, head all-reduce
launched from the ConvDet gradient hook, rest reduced afterwards) against the same step on the full batch in one
process.  Prints the largest gradient difference and the step time."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from squeezedet_pytorch_b200 import dist as sdist, model as M, synth, targets, config
import torch.distributed as dist

rank, world, local = sdist.init_from_env()
dev = torch.device("cuda", local)
shp = synth.KITTI
cfg = config.make_config(shp, device=str(dev), dropout_prob=0.0)
torch.manual_seed(0)
net = M.SqueezeDetWithLoss(cfg).to(dev)
with torch.no_grad():
    w, b = synth.convdet_params(shp, 3)
    net.base.convdet.weight.copy_(torch.from_numpy(w)); net.base.convdet.bias.copy_(torch.from_numpy(b))
net.train()
B = 4 * world
g = torch.Generator().manual_seed(5)
img = torch.randn(B, 3, *shp.input_hw, generator=g).to(dev)
torch.cuda.set_device(dev)
matcher = targets.AnchorMatcher(cfg.anchors, shp.num_classes, device=dev)
cls_l, box_l = zip(*[synth.gt_boxes(shp, 50 + i) for i in range(B)])
gt = matcher.dense_targets(*matcher.pack(list(box_l), list(cls_l)))
batch = {"image": img, "gt": gt}
# single-process reference gradients on the full batch
loss, _ = net(batch)
loss.mean().backward()
want = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
for p in net.parameters():
    p.grad = None
bucket = sdist.bucket_for(net)
mine = sdist.shard_batch(batch, rank, world)
for _ in range(3):
    l, stats = sdist.train_step(net, mine, bucket)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(10):
    l, stats = sdist.train_step(net, mine, bucket)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
worst = 0.0
for n, p in net.named_parameters():
    d = (p.grad - want[n]).abs().max().item() / (want[n].abs().max().item() + 1e-12)
    worst = max(worst, d)
print("rank %d/%d: early segment %d of %d floats, worst relative gradient difference %.2e, step %.2f ms, loss %.4f"
      % (rank, world, bucket.early_numel, bucket.flat.numel(), worst, dt * 1e3, float(l)))
assert worst < 2e-3, worst
dist.destroy_process_group()
