"""Dev tool / measurement: the test-set metrics pass (rbl_test_metrics) on a B200 box.

    python scripts/bench_metrics.py [n] [d]        -> one JSON line

GPU: CUDA events around `reps` launches on the current stream, test set (n x d fp64, default 400k x 1000 = 3.2 GB,
larger than L2) resident in HBM; achieved = n*d*8 algorithmic bytes / launch time against the measured HBM copy
bandwidth (MEASURED_PEAKS.json, fallback 6550.7).  CPU: the oracle restatement of calculate_accuracy +
calculate_statistics (what the reference computes with numpy) on a bounded sample of the same rows."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200.metrics import DeviceTestSet, statistics_from_counts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
gen = torch.Generator(device="cuda").manual_seed(17)
X = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=gen)
w = torch.randn(d, dtype=torch.float64, device="cuda", generator=gen) / d ** 0.5
y = torch.where(X @ w + 0.3 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen) > 0, 1.0, -1.0)
grp = (torch.rand(n, device="cuda", generator=gen) < 0.3).to(torch.int32)
ts = DeviceTestSet(X, y, group=grp)
for _ in range(3):
    ts.launch(w)
torch.cuda.synchronize()
reps = 20
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ts.launch(w)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
t0 = time.perf_counter()
c = ts.counts(w)                    # the user-facing call: launch + 128-byte read-back
e2e_ms = (time.perf_counter() - t0) * 1e3
try:
    peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
except Exception:  # noqa: BLE001
    peak, src = 6550.7, "fallback"
gbs = n * d * 8 / (ms * 1e-3) / 1e9
# CPU sample (oracle = the reference's numpy arithmetic), all host threads numpy's BLAS uses
from oracle import rbl_oracle as O
ns = min(n, 100_000)
Xs, ys, gs, wh = X[:ns].cpu().numpy(), y[:ns].cpu().numpy(), grp[:ns].cpu().numpy(), w.cpu().numpy()
t0 = time.perf_counter()
acc = O.calculate_accuracy(wh, Xs, ys)
st = O.calculate_statistics(wh, Xs, ys, gs)
cpu_s = time.perf_counter() - t0
cs = DeviceTestSet(Xs, ys, group=gs).counts(wh)
print(json.dumps({
    "kernel": "metrics_kernel", "workload": f"test set {n}x{d} fp64, accuracy + group confusion + Theil sums",
    "ms_per_pass": ms, "rows_per_s": n / (ms * 1e-3), "e2e_ms_counts_call": e2e_ms,
    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "peak_source": src, "unit": "GB/s", "frac": gbs / peak,
                 "algorithmic_bytes": n * d * 8},
    "cpu_baseline": {"kind": "port", "sample": f"first {ns} rows, calculate_accuracy + calculate_statistics",
                     "rows_per_s": ns / cpu_s, "cores": os.cpu_count()},
    "check": {"accuracy_gpu": cs[0] / cs[1], "accuracy_cpu": acc,
              "max_abs_stat_diff": float(np.nanmax(np.abs(np.array(statistics_from_counts(cs)) - np.array(st))))},
}))
