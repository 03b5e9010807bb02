#!/usr/bin/env python
"""bench.py — ADMM iterations/sec for SRM (superquantile q=0.8, BCE, l1_reg=0.01) on synthetic
n = 1M x d = 1000 fp64, the configuration BASELINE.json's metric is quoted on (configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1: rows sharded)

A "step" is one ADMM iteration (z-step: margins, radix sort, PAV prox, scatter; w-step: FISTA to its
tolerance; dual update + stop test) of one solve started from the reference's initial state; W
warm-up iterations are iterations 0..W-1 of that solve, the K timed ones are W..W+K-1.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "admm-for-rank-based-loss_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "admm_iters_per_sec"
UNIT = "iterations/s"
Q, L1_REG, TOL = 0.8, 0.01, 1e-6
BLOCK_ROWS = 125_000  # data is generated in fixed 125k-row blocks so every --gpus N solves the SAME problem


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs[0..4]; c2 (the configuration the metric is quoted on) is the default "
                         "and the line the driver records; the others are bench_configs.py")
    ap.add_argument("--n", type=int, default=0, help="rows (0 = the configuration's own)")
    ap.add_argument("--d", type=int, default=1000)
    ap.add_argument("--storage", default="fp64", choices=["fp64", "fp32"],
                    help="c2: fp32 = the OPTIONAL mode that keeps D in float32 in HBM (fp64 arithmetic); a second "
                         "line, never the headline")
    ap.add_argument("--loss", default=None, help="c4: hinge (default, run_AoRR_ratio.py) or binary_cross_entropy")
    ap.add_argument("--instances-per-gpu", type=int, default=32, help="c5")
    ap.add_argument("--batch-mode", default=None, help="c5: gram | stream (default: auto)")
    ap.add_argument("--share-design", action="store_true",
                    help="c4: the solves of the sweep share one device-resident D / G (ADMMmethod(_share=first))")
    ap.add_argument("--no-solve", action="store_true", help="skip the run-to-tolerance tail")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-input e2e variant")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg (development runs)")
    a = ap.parse_args()
    if a.config == "c2" and not a.n:
        a.n = 1_000_000
    return a


def planted_wstar(d):
    rng = np.random.default_rng(17)
    w = np.zeros(d)
    w[:10] = rng.normal(size=min(10, d))
    return w


def gen_rows_numpy(lo, hi, d):
    """CPU twin of the device generator for the reference arm / cpu_baseline sample (same recipe:
    X ~ N(0,1), y = sign(X w* + 0.1 eps); not the same random stream as the device blocks)."""
    rng = np.random.default_rng(1000 + lo)
    X = rng.standard_normal(size=(hi - lo, d))
    y = np.sign(X @ planted_wstar(d) + 0.1 * rng.standard_normal(hi - lo))
    y[y == 0] = 1.0
    return X, y.reshape(-1, 1)


def gen_rows_device(torch, dev, lo, hi, n, d, pin=True):
    """Rows [lo, hi) of THE bench data set, generated on the device in fixed 125k-row blocks with a
    per-block seed — every --gpus N and both --impl arms get bit-identical rows — and returned as host tensors
    (pinned: the e2e leg uploads from them inside its timed region)."""
    wstar = torch.from_numpy(planted_wstar(d)).to(dev)
    X_host = torch.empty((hi - lo, d), dtype=torch.float64, pin_memory=pin)
    y_host = torch.empty(hi - lo, dtype=torch.float64, pin_memory=pin)
    for blk in range(lo // BLOCK_ROWS, (hi + BLOCK_ROWS - 1) // BLOCK_ROWS):
        b_lo, b_hi = blk * BLOCK_ROWS, min(n, (blk + 1) * BLOCK_ROWS)
        g = torch.Generator(device=dev)
        g.manual_seed(17 + blk)
        Xb = torch.randn(b_hi - b_lo, d, generator=g, dtype=torch.float64, device=dev)
        eb = torch.randn(b_hi - b_lo, generator=g, dtype=torch.float64, device=dev)
        yb = torch.sign(Xb @ wstar + 0.1 * eb)
        yb[yb == 0] = 1.0
        s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
        X_host[s_lo - lo: s_hi - lo].copy_(Xb[s_lo - b_lo: s_hi - b_lo])
        y_host[s_lo - lo: s_hi - lo].copy_(yb[s_lo - b_lo: s_hi - b_lo])
        del Xb, eb, yb
    torch.cuda.synchronize(dev)
    return X_host, y_host


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


def use_all_host_threads():
    """all host cores for BLAS, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 per rank)"""
    cores = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:  # noqa: BLE001
        pass
    try:
        from threadpoolctl import threadpool_limits
        use_all_host_threads._limits = threadpool_limits(limits=cores)  # kept alive: applies process-wide
    except Exception:  # noqa: BLE001
        pass
    return cores


def cpu_reference_run(X, y, W, K, after_step=None):
    """The oracle port (numpy/BLAS matvecs with all host threads, C stack-PAV, numpy FISTA fp64: the reference's
    algorithm with its O(n^2) Python PAV replaced by an exact O(n) one) on the FULL data set the GPU arm solves:
    iterations 0..W-1 untimed, W..W+K-1 timed one by one (host clock around each oracle step only, so a lockstep
    GPU comparison in `after_step` is not charged).  Returns (iterations/s, seconds, oracle)."""
    from oracle import rbl_oracle as O

    o = O.OracleADMM(X, y, "superquantile", "binary_cross_entropy", l1_reg=L1_REG, args=[Q], max_iter=100_000,
                     tol=TOL)
    dt = 0.0
    for i in range(W + K):
        t0 = time.perf_counter()
        o.step()
        if i >= W:
            dt += time.perf_counter() - t0
        if after_step is not None:
            after_step(i, o)
    return K / dt, dt, o


CPU_SAMPLE = ("oracle port (numpy/BLAS matvecs on all host threads, C stack-PAV, numpy FISTA fp64) on the SAME {n} x {d} "
              "rows the GPU arm solves (device-generated blocks copied to the host), ADMM iterations {w}..{last} of "
              "one solve from the reference's initial state ({dt:.1f} s of CPU work), no scaling")


def run_reference(args, rank):
    """--impl reference: the CPU implementation of the path on this box's host cores, same data / config / steps."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    import torch

    n, d, W, K = args.n, args.d, args.warmup, args.steps
    if torch.cuda.is_available():
        # data only: the same device-generated blocks as the GPU arm, copied to (pageable) host memory
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        Xh, yh = gen_rows_device(torch, dev, 0, n, n, d, pin=False)
        X, y, same = Xh.numpy(), yh.numpy().reshape(-1, 1), True
        torch.cuda.empty_cache()
    else:
        X, y = gen_rows_numpy(0, n, d)
        same = False
    v, dt, o = cpu_reference_run(X, y, W, K)
    sample = CPU_SAMPLE.format(n=n, d=d, w=W, last=W + K - 1, dt=dt)
    if not same:
        sample += " [no CUDA device here: numpy twin of the recipe, NOT the same random stream]"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, d),
                   "same_data_as_gpu_arm": same,
                   "note": "the reference is pure Python with an O(n^2) sweep-PAV and cannot run this n (days per "
                           "z-step); its restated CPU port (exact stack PAV, same FISTA) is timed instead"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "state_after_timed_steps": {"primal": o.primal, "dual": o.dual, "rho": float(o.rho),
                                    "objective": o.objective(), "nnz_w": int(np.count_nonzero(o.w))},
    }), flush=True)


def cpu_baseline_with_parity(torch, ADMMmethod, Optimizer, X_host, y_host, kw, W, K, n, d, quiet, timed_solver):
    """cpu_baseline (N = 1, rank 0): the oracle port on the same rows, same W / K, no scaling — and, from the same
    run, full-size parity: a second GPU solver built from the same host arrays is stepped ONE iteration at a time
    (the per-iteration path replays the same graph as the native loop, bit for bit) in lockstep with the free-running
    oracle; rel |w|, rel |z| are recorded after every iteration, both solvers run on without any re-synchronisation.
    north_star tolerance: 1e-9 relative on iterates and objective (fp64)."""
    cores = use_all_host_threads()
    Xn, yn = X_host.numpy(), y_host.numpy().reshape(-1, 1)
    with quiet:
        par = ADMMmethod(Xn, yn, **kw)
    if kw.get("_storage") == "fp32":
        # the fp32 mode IS the fp64 algorithm on the design matrix rounded to float32: the oracle gets that matrix
        Xn = Xn.astype(np.float32).astype(np.float64)
    rows = []

    def rel(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    def after_step(i, o):
        with quiet:
            Optimizer.main_loop(par, i, 0.0, False)
        z_gpu = par.engine.z.cpu().numpy()
        fi = getattr(o, "last_fista_info", (None, None, None))
        rows.append({"iteration": i, "timed": i >= W, "rel_w": rel(par.w.reshape(-1), o.w), "rel_z": rel(z_gpu, o.z),
                     "rel_rho": abs(float(par.rho) - float(o.rho)) / float(o.rho),
                     "rel_primal": abs(par.primal_feasibility - o.primal) / max(o.primal, 1e-300),
                     "fista_iters_cpu": fi[0]})

    v, dt, o = cpu_reference_run(Xn, yn, W, K, after_step)
    with quiet:
        obj_gpu = par.objective.get_arrogate_loss(torch.from_numpy(par.w).double())
    obj_cpu = o.objective()
    # the timed solver has moved on to the stop test; its state after W + K iterations was the lockstep solver's
    # (same inputs, same kernels: test_native_loop_equals_per_iteration_loop) — checked here on the final objective
    tol = 1e-9
    timed = [r for r in rows if r["timed"]]
    parity = {"tolerance": tol, "iterations_compared": len(rows),
              "max_rel_w": max(r["rel_w"] for r in rows), "max_rel_z": max(r["rel_z"] for r in rows),
              "max_rel_w_timed": max(r["rel_w"] for r in timed), "max_rel_z_timed": max(r["rel_z"] for r in timed),
              "max_rel_rho": max(r["rel_rho"] for r in rows), "max_rel_primal": max(r["rel_primal"] for r in rows),
              "objective_gpu": obj_gpu, "objective_cpu": obj_cpu,
              "rel_objective": abs(obj_gpu - obj_cpu) / abs(obj_cpu),
              "per_iteration": [{k: (float("%.3e" % v) if isinstance(v, float) else v) for k, v in r.items()}
                                for r in rows],
              "how": "free-running lockstep of a GPU solver (per-iteration path) and the CPU oracle on the same "
                     "n x d rows, no re-synchronisation; every iteration compared"}
    parity["ok"] = bool(parity["max_rel_w"] <= tol and parity["max_rel_z"] <= tol and parity["rel_objective"] <= tol)
    par.engine.close()
    cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": CPU_SAMPLE.format(n=n, d=d, w=W, last=W + K - 1, dt=dt)}
    return cpu, parity


def workload_name(n, d):
    return f"SRM superquantile(q={Q}) BCE l1_reg={L1_REG} ADMM, n={n} d={d} fp64 (BASELINE configs[1])"


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config != "c2":
        import bench_configs

        me = sys.modules[__name__]
        if args.impl == "reference":
            if args.config == "c5":
                raise SystemExit("--impl reference --config c5: see the cpu_baseline block of the c5 line")
            bench_configs.run_reference(args, me)
        elif args.config == "c5":
            bench_configs.run_c5(args, me)
        else:
            bench_configs.run(args, me)
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from rbl_b200 import _cabi
    from rbl_b200.engine import shard_bounds
    from src.optim.algorithms import ADMMmethod, Optimizer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, K, W = args.n, args.d, args.steps, args.warmup
    lo, hi = shard_bounds(n, world, rank)

    # ---- synthetic planted data, generated on the device in fixed blocks, parked in pinned host
    # memory: the e2e leg uploads from there inside its timed region ------------------------------------------
    X_host, y_host = gen_rows_device(torch, dev, lo, hi, n, d, pin=True)
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=L1_REG, args=[Q],
              max_iter=100_000, tol=TOL)
    if args.storage == "fp32":
        kw["_storage"] = "fp32"
    shard = dict(row_lo=lo, n_global=n) if world > 1 else {}

    # ---- process warm-up (untimed): a tiny solve of the same kind, so that the one-off costs of a fresh process —
    # librbl_b200's module load on its first call, lazy loading of each kernel on its first launch, function
    # attributes — are not charged to the end-to-end figure (the CUDA context itself is already up: the data above
    # was generated on the device).  Same d, hence the same kernel instantiations; 4096 rows.
    rng_w = np.random.default_rng(5)
    Xw = rng_w.standard_normal((4096, d))
    yw = np.sign(Xw @ planted_wstar(d) + 0.1 * rng_w.standard_normal(4096)).reshape(-1, 1)
    yw[yw == 0] = 1.0
    t_cold0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        warm = ADMMmethod(Xw, yw, **{**kw, "max_iter": 8})
        warm.advance(0, 8)
    torch.cuda.synchronize()
    t_cold = time.perf_counter() - t_cold0   # the one-off cost a fresh process pays on its first solve
    warm.engine.close()
    del warm, Xw, yw
    torch.cuda.synchronize()

    # ---- upload + build (e2e part 1): pinned host -> HBM, D = -y (.) X -------------------------------
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    solver = ADMMmethod(X_host.numpy(), y_host.numpy().reshape(-1, 1), _shard=shard, **kw)
    ev[1].record()
    barrier()
    t_upload = ev[0].elapsed_time(ev[1]) / 1e3
    h2d_bytes = X_host.numel() * 8 + y_host.numel() * 8
    launches0 = solver.engine.launches

    quiet = contextlib.redirect_stdout(io.StringIO())
    done = False
    w0_ev, w1_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0_ev.record()
    with quiet:
        _, done = solver.advance(0, W)
    w1_ev.record()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches1 = solver.engine.launches
    passes1 = solver.engine.fista_stats["passes"]
    dpasses1 = solver.engine.fista_stats["d_passes"]
    ev[2].record()
    with quiet:
        it_next, done2 = solver.advance(W, K)
    done = done or done2
    ev[3].record()
    barrier()
    clocks = None
    if rank == 0 and (args.no_solve or done):
        clocks = sampler.stop()
    t_steps = ev[2].elapsed_time(ev[3]) / 1e3
    t_warm = w0_ev.elapsed_time(w1_ev) / 1e3
    launches_timed = solver.engine.launches - launches1
    passes_timed = solver.engine.fista_stats["passes"] - passes1
    dpasses_timed = solver.engine.fista_stats["d_passes"] - dpasses1
    if world > 1:
        t = torch.tensor([t_steps, t_upload], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_steps, t_upload = float(t[0]), float(t[1])
    value = K / t_steps
    e2e_iters = K
    e2e_value = K / (t_steps + t_upload)  # replaced below by the whole-solve figure when the solve is run

    # ---- roofline of the D-reading kernels, each timed alone on its stream ------------------------------
    eng = solver.engine
    reps = 20
    peak, peak_kind = measured_peak_gbs()
    nl = hi - lo

    def time_kernel(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(reps):
            fn()
        p1.record()
        torch.cuda.synchronize()
        return p0.elapsed_time(p1) / 1e3 / reps

    xd = eng.w.clone()
    t_pass = time_kernel(lambda: _cabi.check(eng.lib.rbl_fused_pass(
        eng.h, eng.D.data_ptr(), xd.data_ptr(), eng.b.data_ptr(), eng.r.data_ptr(), eng.red.data_ptr(),
        eng._stream())))
    esz = (8 if args.storage == 'fp64' else 4)  # bytes per stored element of D
    alg_bytes = nl * d * esz + (2 * nl + 2 * d) * 8  # D once; b in, r out; x in, g out  
    achieved = alg_bytes / t_pass / 1e9
    gather = None
    if eng.w_mode == "gram" and eng.active_dense_frac > 0:
        import ctypes
        eng.z_step(solver.rho)  # leaves the active-row list of the current state
        cnt = ctypes.c_int32(0)
        _cabi.check(eng.lib.rbl_active_count(eng.h, ctypes.byref(cnt), eng._stream()))
        t_g = time_kernel(lambda: _cabi.check(eng.lib.rbl_gather_only(eng.h, eng.D.data_ptr(), eng._stream())))
        g_bytes = cnt.value * (d * esz + 12) + 2 * d * 8  # active rows of D + (row, delta) list; g out
        gather = {"kernel": "rbl_gather_kernel (g = sum over ACTIVE rows of delta_i D_i, per-row TMA bulk copies)",
                  "active_rows": cnt.value, "active_fraction": cnt.value / nl, "launch_ms": 1e3 * t_g,
                  "algorithmic_bytes_per_launch": g_bytes, "achieved": g_bytes / t_g / 1e9,
                  "frac": g_bytes / t_g / 1e9 / peak}
    # z-step alone (sort + PAV + scatter): 68 n bytes algorithmic
    torch.cuda.synchronize()
    z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    z0.record()
    for _ in range(5):
        eng.z_step(solver.rho)
    z1.record()
    torch.cuda.synchronize()
    t_z = z0.elapsed_time(z1) / 1e3 / 5

    # ---- optional tail: keep iterating the same solve to the 1e-6 stop test ---------------------------
    solve = None
    if not args.no_solve:
        eng._r_matches_w = False
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        it2 = W + K
        with quiet:
            if not done:
                it2, done = solver.advance(W + K, 1000 - (W + K))
        s1.record()
        barrier()
        if rank == 0 and clocks is None:
            # the K timed steps last ~20 ms: keep sampling through the continuation of the SAME solve (same
            # kernels, same load) so that the 20 ms nvidia-smi period yields several samples under load
            clocks = sampler.stop()
            clocks["window"] = "timed steps + continuation of the same solve to the stop test"
        t_tail = s0.elapsed_time(s1) / 1e3
        with quiet:
            obj = solver.objective.get_arrogate_loss(torch.from_numpy(solver.w).double())
        solve = {"iterations": it2, "converged": bool(done),
                 "time_to_1e-6_s": round(t_warm + t_steps + t_tail, 3),
                 "time_to_1e-6_s_incl_upload": round(t_upload + t_warm + t_steps + t_tail, 3), "objective": obj, "nnz_w": int(np.count_nonzero(solver.w)),
                 "phases_s": {"warmup_iterations_0_to_%d" % (W - 1): round(t_warm, 4),
                              "timed_%d_iterations" % K: round(t_steps, 4),
                              "continuation_to_stop_test": round(t_tail, 4)},
                 "primal": solver.primal_feasibility, "dual": solver.dual_feasibility,
                 "fista_passes_total": eng.fista_stats["passes"], "fista_calls": eng.fista_stats["calls"]}
        # degenerate-benchmark guard: the solve must do real work
        solve["non_degenerate"] = bool(obj < np.log(2.0) and solve["nnz_w"] > 0)
        t_e2e = t_upload + t_warm + t_steps + t_tail
        if world > 1:
            tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt[0])
        e2e_value, e2e_iters = it2 / t_e2e, it2

    # ---- the same call on a PAGEABLE numpy array (what a user's array is; X_host above is pinned) ------------------
    e2e_pageable = None
    if solve is not None and not args.no_pageable:
        X_page = np.array(X_host.numpy())          # ordinary (pageable) copy, made outside the timed region
        y_page = np.array(y_host.numpy()).reshape(-1, 1)
        reps = []
        for rep in range(2):
            barrier()
            g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            g0.record()
            with quiet:
                s2 = ADMMmethod(X_page, y_page, _shard=shard, **kw)
                g1.record()
                it_p, done_p = s2.advance(0, 1000)
            g2.record()
            barrier()
            tp_build, tp_all = g0.elapsed_time(g1) / 1e3, g0.elapsed_time(g2) / 1e3
            if world > 1:
                tt = torch.tensor([tp_build, tp_all], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                tp_build, tp_all = float(tt[0]), float(tt[1])
            reps.append({"value": it_p / tp_all, "unit": UNIT, "iterations": it_p, "converged": bool(done_p),
                         "upload_and_build_s": tp_build, "total_s": tp_all,
                         "upload_path": getattr(s2.engine, "upload_path", None),
                         "same_result": bool(np.array_equal(s2.w, solver.w))})
            if rep == 0:
                s2.engine.close()
        # the array is seconds old at its first use (the OS is still placing its pages: 3-5x slower on these boxes,
        # scripts/upload_time.py shows the same for any fresh copy); a caller's array has usually been around longer
        e2e_pageable = dict(reps[1], first_use_of_a_fresh_array=reps[0])
        s2.engine.close()
        del s2, X_page, y_page

    def shutdown():
        # release the captured iteration graph (it holds NCCL work) before the process group goes away
        solver.engine.close()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        shutdown()
        return

    cpu, parity = None, None
    if world == 1 and not args.no_cpu:
        cpu, parity = cpu_baseline_with_parity(torch, ADMMmethod, Optimizer, X_host, y_host, kw, W, K, n, d, quiet,
                                               solver)

    stream_pass = {"kernel": "rbl_pass_kernel (fused r = b - D x, ||r||^2, D^T r over ALL rows)",
                   "achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_launch": alg_bytes,
                   "launch_ms": 1e3 * t_pass, "launches_in_timed_region": dpasses_timed,
                   "share_of_step": dpasses_timed * t_pass / t_steps}
    if gather is not None and eng.active_stats["gathered"] > 0:
        # the D-reading kernel of the step is the active-row gather; the streaming pass is kept for reference
        g_launches = K  # one gradient pass per ADMM iteration
        traffic, traffic_note = None, "no ncu capture on file"
        try:
            if args.storage != "fp64":
                raise RuntimeError("the ncu capture on file is of the fp64 kernel")
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_gather_traffic.json")))
            per_row = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["active_rows"]
            traffic = per_row * gather["active_rows"]
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this "
                            "kernel (%d active rows: %.4f GB), scaled per active row to this launch; "
                            "profiles/r01_gather_traffic.json" % (tj["active_rows"],
                                                                  (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / 1e9))
        except Exception:  # noqa: BLE001
            pass
        roofline = {"bound": "hbm", "kernel": gather["kernel"], "achieved": gather["achieved"], "peak": peak,
                    "peak_kind": peak_kind, "unit": "GB/s", "frac": gather["frac"], "traffic": traffic,
                    "traffic_note": traffic_note,
                    "algorithmic_bytes_per_launch": gather["algorithmic_bytes_per_launch"],
                    "launch_ms": gather["launch_ms"], "active_rows": gather["active_rows"],
                    "active_fraction": gather["active_fraction"],
                    "share_of_step": g_launches * gather["launch_ms"] * 1e-3 / t_steps,
                    "stream_pass": stream_pass}
    else:
        roofline = {"bound": "hbm", "kernel": stream_pass["kernel"], "achieved": achieved, "peak": peak,
                    "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": 1e3 * t_pass,
                    "share_of_step": dpasses_timed * t_pass / t_steps}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_steps / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.storage == "fp64" else "f64 arithmetic on f32-stored D (optional mode)",
        "data": "synthetic",
        "config": {"workload": workload_name(n, d) + ("" if args.storage == "fp64" else
                                                     " — OPTIONAL fp32 storage of D, not the headline"),
                   "storage": args.storage,
                   "rows_per_gpu": hi - lo, "parallelism": f"rows sharded x{world}" if world > 1 else "single GPU",
                   "timed_iterations": f"{W}..{W + K - 1} of one solve from the reference's initial state",
                   "l2_flush": "none needed: every D pass streams %.1f GB per GPU, far above the 126 MB L2"
                               % ((hi - lo) * d * 8 / 1e9),
                   "w_mode": solver.engine.w_mode,
                   "w_mode_note": "gram: FISTA trials sweep G = D^T D (d x d, L2-resident), D is read twice per "
                                  "ADMM iteration (D^T b, D w); stream: every FISTA trial is a pass over D "
                                  "(RBL_W_MODE=stream)",
                   "fista_trials_in_timed_region": passes_timed,
                   "d_passes_in_timed_region": dpasses_timed,
                   "gram_build_s": solver.engine.gram_build_s,
                   "build_times_s": {k: round(v, 4) for k, v in solver.engine.build_times.items()},
                   "active_rows": {**eng.active_stats, "note": "gradient pass reads only rows with z != m "
                                   "(b - D w = z - m is exactly 0 elsewhere); rows = total rows read over `calls`"},
                   "dual_pass": {**eng.dual_stats, "note": "D w reads only the sectors touched by nnz(w) when "
                                 "w is sparse"},
                   "pass_tiles": {k: solver.engine.info[k] for k in ("pass_grid", "rows_per_tile", "pass_stages")}},
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": h2d_bytes / (e2e_iters if solve else K),
                "d2h_bytes_per_step": (d + 16) * 8,
                "note": ("the call a user makes: ADMMmethod(X, y, ...) on HOST numpy arrays (pinned -> HBM upload "
                         "of X and y, D = -y*X, G = D^T D) followed by the ADMM loop to the 1e-6 stop test, w and "
                         "the residuals read back to the host every iteration; value = all %d iterations of that "
                         "solve / (upload + build + solve) device time" % e2e_iters) if solve else
                        "K iterations through ADMMmethod plus the whole host -> HBM upload and build charged to "
                        "the K steps (--no-solve)",
                "upload_and_build_s": t_upload,
                "input": "pinned host memory (best case)",
                "pageable_input": e2e_pageable,
                "cold_first_call_s": {"value": t_cold,
                                      "what": "wall time of the untimed warm-up solve (4096 x d, 8 iterations) that "
                                              "a fresh process runs first: library load, lazy kernel loading, "
                                              "function attributes, first allocations — paid once per process, "
                                              "NOT included in the e2e figures"}},
        "gpu_launches": launches_timed,
        "clocks": clocks,
        "roofline": roofline,
        "zstep": {"ms": 1e3 * t_z, "keys_per_s": n / t_z, "algorithmic_bytes": 68 * n,
                  "frac_of_hbm_peak": 68 * n / t_z / 1e9 / peak},
        "cpu_baseline": cpu,
        "parity": parity,
        "solve": solve,
    }
    print(json.dumps(out), flush=True)
    shutdown()


if __name__ == "__main__":
    main()
