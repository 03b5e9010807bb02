"""Dev tool / measurement: device-side standardisation (rbl_standardize_columns) and split gather on a B200 box.

    python scripts/bench_ingest.py [n] [d]    -> one JSON line

CUDA events around `reps` calls on a matrix resident in HBM (default 1 M x 1000 fp64 = 8 GB).  Algorithmic bytes of a
standardisation: 3 reads + 1 write of n*d*8; of a split: 1 read + 1 write of the selected rows.  CPU: scikit-learn's
`preprocessing.scale` on a bounded row sample (what load_data.py:115 runs)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import ingest

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
gen = torch.Generator(device="cuda").manual_seed(17)
X = torch.randn((n, d + (d & 1)), dtype=torch.float64, device="cuda", generator=gen) * 3.0 + 1.0
def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms_std = timeit(lambda: ingest.standardize_(X, d), 5)
idx = torch.randperm(n, device="cuda", generator=gen)[: int(0.6 * n)]
ms_split = timeit(lambda: ingest.split_rows(X, idx, d), 5)
try:
    peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
except Exception:  # noqa: BLE001
    peak, src = 6550.7, "fallback"
b_std, b_split = 4 * n * d * 8, 2 * idx.numel() * d * 8
from sklearn import preprocessing
ns = min(n, 100_000)
Xs = X[:ns, :d].cpu().numpy()
t0 = time.perf_counter(); preprocessing.scale(Xs); cpu_s = time.perf_counter() - t0
print(json.dumps({
    "workload": f"{n}x{d} fp64 resident in HBM",
    "standardize": {"ms": ms_std, "algorithmic_bytes": b_std, "achieved_gbs": b_std / ms_std / 1e6,
                    "frac_of_hbm_peak": b_std / ms_std / 1e6 / peak, "rows_per_s": n / ms_std * 1e3},
    "split_gather_60pct": {"ms": ms_split, "algorithmic_bytes": b_split, "achieved_gbs": b_split / ms_split / 1e6,
                           "frac_of_hbm_peak": b_split / ms_split / 1e6 / peak},
    "peak_gbs": peak, "peak_source": src,
    "cpu_baseline": {"kind": "sklearn.preprocessing.scale", "sample_rows": ns, "rows_per_s": ns / cpu_s,
                     "cores": os.cpu_count()},
}))
