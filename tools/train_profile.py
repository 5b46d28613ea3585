"""Kernel-level breakdown of the training step bench.py times as `train_step` (BASELINE configs[3], one GPU): runs the eager
step under torch.profiler and prints every CUDA kernel / memcpy with its time per step.  usage: python tools/train_profile.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import config as sconfig, dist as sdist, model as smodel, synth, targets as stargets
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
shp, B = synth.KITTI, 20
cfg = sconfig.make_config(shp, device=dev)
net = smodel.SqueezeDetWithLoss(cfg).to(dev)
w_np, b_np = synth.convdet_params(shp, 4321)
with torch.no_grad():
    net.base.convdet.weight.copy_(torch.from_numpy(w_np))
    net.base.convdet.bias.copy_(torch.from_numpy(b_np))
net.train()


class HeadWithLoss(torch.nn.Module):
    def __init__(self, full):
        super().__init__()
        self.base, self.loss = full.base, full.loss

    def forward(self, batch):
        return self.loss(self.base.head(batch["features"]), batch["gt"])


head = HeadWithLoss(net)
bucket = sdist.bucket_for(net)
matcher = stargets.AnchorMatcher(cfg.anchors, shp.num_classes, device=dev)
cls_l, box_l = zip(*[synth.gt_boxes(shp, 100 + i) for i in range(B)])
gt_packed = matcher.pack(list(box_l), list(cls_l))
tfeat = torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev)).requires_grad_(True)


class Sink(torch.autograd.Function):   # the backbone's backward would take the gradient from here (see bench.py)
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return None


def it():
    tfeat.grad = None
    gt = matcher.dense_targets(*gt_packed)
    return sdist.train_step(head, {"features": Sink.apply(tfeat) if "--leaf" not in sys.argv else tfeat, "gt": gt}, bucket)


for _ in range(5):
    it()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        it()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"device time per step {tot / N:.1f} us over {sum(e.count for e in rows) / N:.0f} launches")
for e in rows[:40]:
    print(f"{e.device_time_total / N:9.1f} us  x{e.count / N:4.1f}  {e.key[:110]}")
