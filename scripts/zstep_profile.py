"""Time the z-step components separately with CUDA events (dev tool; run on the B200 box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi  # noqa: E402
from rbl_b200 import spectra  # noqa: E402
from rbl_b200.engine import AdmmEngine  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rng = np.random.default_rng(0)
    sig = spectra.get_superquantile_weights(n, 0.8)
    e = AdmmEngine(np.zeros((n, 2)), np.ones(n), "binary_cross_entropy", sig)
    lib, h, s = e.lib, e.h, e._stream
    m = e.vec(rng.normal(size=n))
    t_sort = timeit(lambda: _cabi.check(lib.rbl_sort_margins(h, m.data_ptr(), e.m_sorted.data_ptr(),
                                                             e.perm.data_ptr(), s())))
    print(f"n={n}: sort {t_sort:.1f} us  ({n / t_sort:.1f} Mkeys/s)")
    for rho, scale, tag in [(1e-5, 1e-3, "giant block (iteration 0)"), (1e-3, 1.0, "mixed"), (1.0, 3.0, "few merges")]:
        ms = e.vec(np.sort(rng.normal(size=n) * scale))
        t = timeit(lambda: _cabi.check(lib.rbl_pav_prox(h, 0, ms.data_ptr(), rho, e.z_sorted.data_ptr(), s())))
        z = e.z_sorted.cpu().numpy()
        print(f"   pav rho={rho:g} [{tag}]: {t:.1f} us, distinct values {len(np.unique(z))}")
    t = timeit(lambda: _cabi.check(lib.rbl_scatter_z(h, e.z_sorted.data_ptr(), e.perm.data_ptr(), 0, 0.0,
                                                     e.lam.data_ptr(), 1e-3, e.z.data_ptr(), e.b.data_ptr(), s())))
    print(f"   scatter {t:.1f} us")
    t = timeit(lambda: _cabi.check(lib.rbl_margins(h, e.Dw.data_ptr(), e.lam.data_ptr(), 1e-3, e.m.data_ptr(), s())))
    print(f"   margins {t:.1f} us")
    t = timeit(lambda: e.z_step(1e-3))
    print(f"   whole z_step (python) {t:.1f} us")


if __name__ == "__main__":
    main()
