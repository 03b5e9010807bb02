"""Dev tool: phase timing inside the persistent radix sort (globaltimer stamps of CTA 0), B200 box."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from rbl_b200.engine import AdmmEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
e = AdmmEngine(np.zeros((n, 2)), np.ones(n), "binary_cross_entropy", np.ones(n) / n)
m = e.vec(np.random.default_rng(0).normal(size=n))
st = torch.zeros(64, dtype=torch.int64, device=e.device)
_cabi.check(e.lib.rbl_sort_debug(e.h, st.data_ptr()))
for _ in range(3):
    _cabi.check(e.lib.rbl_sort_margins(e.h, m.data_ptr(), e.m_sorted.data_ptr(), e.perm.data_ptr(), e._stream()))
torch.cuda.synchronize()
t = st.cpu().numpy()[:48].reshape(8, 6).astype(np.float64)
names = ["load+rank+count", "barrier1", "prefix over CTAs", "reorder+scatter", "barrier2"]
d = np.diff(t, axis=1) / 1e3
print("per pass (us):")
for i, nm in enumerate(names):
    print(f"  {nm:18s} " + " ".join(f"{x:6.1f}" for x in d[:, i]) + f"   mean {d[:, i].mean():6.1f}")
print(f"  total {(t[7, 5] - t[0, 0]) / 1e3:.1f} us")
