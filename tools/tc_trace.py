"""Debug: per-unit pipeline timeline of CTA 0 of the tcgen05 ConvDet kernel (clock64 stamps)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squeezedet_pytorch_b200 import ops, synth

shp = synth.KITTI
feat = torch.relu(torch.randn((20, 768, 24, 78), device="cuda")).contiguous(memory_format=torch.channels_last)
w, b = synth.convdet_params(shp, 9)
w, b = torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda()
packed = ops.pack_convdet_weights(w)
for _ in range(3):
    ops.convdet_forward(feat, w, b, packed=packed)
trace = torch.zeros((512, 32), dtype=torch.int64, device="cuda")
os.environ["SQD_TC_TRACE"] = hex(trace.data_ptr())
ops.convdet_forward(feat, w, b, packed=packed, check_status=True)
del os.environ["SQD_TC_TRACE"]
t = trace.cpu().numpy()
n = int((t[:, 3] > 0).sum())
t0 = t[0, 0]
names = ["A_tma_issue", "cvt_start", "cvt_done", "mma_start", "mma_issued", "acc_start", "acc_done", "B_tma_issue",
         "c0_computed", "c0_slotfree", "c0_stored", "c1_computed", "c1_slotfree", "c1_stored", "c2_computed", "-"]
print("unit " + " ".join(f"{x:>12s}" for x in names) + "   (cycles since first A issue)")
for i in list(range(0, 6)) + list(range(70, 74)) + list(range(n - 3, n)):
    print(f"{i:4d} " + " ".join(f"{int(v - t0):9d}" for v in t[i][:25]))
d = np.diff(t[:n, 3])
print("units", n, "mean period (mma_start to mma_start)", d.mean(), "median", np.median(d))
print("mean cvt_start - A_issue (TMA latency)", (t[:n, 1] - t[:n, 0]).mean(), " cvt duration", (t[:n, 2] - t[:n, 1]).mean())
print("mean mma_start - cvt_done", (t[:n, 3] - t[:n, 2]).mean(), " mma issue duration", (t[:n, 4] - t[:n, 3]).mean())
print("mean acc_start - mma_issued (MMA drain)", (t[:n, 5] - t[:n, 4]).mean(), " acc duration", (t[:n, 6] - t[:n, 5]).mean())
print("converter dy0: compute", (t[:n, 8] - t[:n, 1]).mean(), "wait slot", (t[:n, 9] - t[:n, 8]).mean(), "store", (t[:n, 10] - t[:n, 9]).mean())
print("converter dy1: compute", (t[:n, 11] - t[:n, 10]).mean(), "wait slot", (t[:n, 12] - t[:n, 11]).mean(), "store", (t[:n, 13] - t[:n, 12]).mean())
print("converter dy2: compute", (t[:n, 14] - t[:n, 13]).mean())
for d in range(3):
    print(f"mma dy{d}: wait A slot", (t[1:n, 16 + 3 * d] - (t[1:n, 3] if d == 0 else t[1:n, 16 + 3 * d - 1])).mean(),
          "wait B", (t[1:n, 17 + 3 * d] - t[1:n, 16 + 3 * d]).mean(), "issue", (t[1:n, 18 + 3 * d] - t[1:n, 17 + 3 * d]).mean())
print("A ready(i,0) - c0_stored(i)", (t[1:n, 16] - t[1:n, 10]).mean(), " slot0 free seen by converter(i+1) - step(i,0) issued", (t[2:n, 9] - t[1:n - 1, 18]).mean())
print("B issue(i, dy0) lead over mma need:", (t[1:n, 17] - t[1:n, 7]).mean())
