"""Dev tool: do the per-instance iteration graphs of the batched Gram mode overlap on the GPU?  Kernel intervals from
CUPTI (torch.profiler) over a few batched steps: busy time (union of intervals), sum of kernel times, wall."""
import os, sys
from collections import defaultdict
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
import bench as B
from rbl_b200.batched import BatchedADMM
from torch.profiler import ProfilerActivity, profile

n, d, inst = 100_000, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
Xh, yh = B.gen_rows_device(torch, dev, 0, n, n, d, pin=True)
regs = np.logspace(-4, 0, 256)[np.linspace(0, 255, inst).round().astype(int)]
b = BatchedADMM(Xh.numpy(), yh.numpy().reshape(-1, 1), "superquantile", "binary_cross_entropy", l1_regs=list(regs),
                args=[0.8], max_iter=10000, tol=1e-6)
for _ in range(8):
    b.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        b.step()
    torch.cuda.synchronize()
iv = []
per = defaultdict(float)
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        t0 = ev.time_range.start
        iv.append((t0, t0 + ev.device_time, ev.name))
        nm = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        per[nm] += ev.device_time
iv.sort()
busy, cur_s, cur_e = 0.0, None, None
for s, e, _ in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
wall = iv[-1][1] - iv[0][0]
tot = sum(e - s for s, e, _ in iv)
print(f"{inst} instances, 4 batched steps: wall {wall / 4e3:.2f} ms/step, GPU busy {busy / 4e3:.2f} ms/step, "
      f"sum of kernel times {tot / 4e3:.2f} ms/step  -> average concurrency {tot / busy:.2f}")
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {k[:60]:60s} {v / 4e3:8.3f} ms/step")
b.close()
