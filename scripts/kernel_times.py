"""Dev tool: warm per-kernel device times of the replayed ADMM iteration (CUPTI through torch.profiler — the kernels
of librbl_b200 launched from the captured CUDA graph are recorded like any other), B200 box.

    python scripts/kernel_times.py [n] [d] [first_iteration] [iterations] [storage]

Unlike the ncu launch list (cold caches, serialised replays) these are the durations inside the running solve."""
import contextlib, io, json, os, sys
from collections import defaultdict
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
import bench as B
from src.optim.algorithms import ADMMmethod

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
first = int(sys.argv[3]) if len(sys.argv) > 3 else 40
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
storage = sys.argv[5] if len(sys.argv) > 5 else "fp64"
cfg = sys.argv[6] if len(sys.argv) > 6 else "c2"      # c2 | c3 | c4 (problem family; n and d as given)
dev = torch.device("cuda", 0)
Xh, yh = B.gen_rows_device(torch, dev, 0, n, n, d, pin=True)
Xn = Xh.numpy()
if cfg == "c3":
    kw = dict(weight_function="ehrm", loss="binary_cross_entropy", B=-5, l2_reg=0.01)
elif cfg == "c4":
    kw = dict(weight_function="aorr", loss="hinge", args=[0.2, 0.8], l2_reg=1e-4)
    Xn = np.hstack((Xn, np.ones((n, 1))))
else:
    kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.01, args=[0.8])
s = ADMMmethod(Xn, yh.numpy().reshape(-1, 1), max_iter=100000, tol=1e-12, _storage=storage, **kw)
quiet = contextlib.redirect_stdout(io.StringIO())
with quiet:
    s.advance(0, first)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0.record()
    with quiet:
        s.advance(first, iters)
    t1.record()
    torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "")
        if name.startswith("void "):
            name = name[5:]
        name = name.split("(")[0]
        a = agg[name]
        a[0] += 1
        a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"# {cfg} n={n} d={d} storage={storage}: iterations {first}..{first + iters - 1}, {ms / iters * 1e3:.1f} us per iteration under the "
      f"profiler, kernels sum {tot / iters:.1f} us per iteration")
print(f"{'kernel':70s} {'per it':>7s} {'mean us':>9s} {'us / it':>9s} {'share':>7s}")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {c / iters:7.2f} {t / c:9.2f} {t / iters:9.2f} {100 * t / tot:6.1f}%")
s.engine.close()
