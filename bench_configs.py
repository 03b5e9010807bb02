"""bench.py --config c1|c3|c4|c5: the other BASELINE.json configurations at their named shapes, same JSON contract as
the headline line (config c2, bench.py itself): device-timed `value`, `e2e` from host arrays, `roofline` of the
D-reading kernel, `cpu_baseline` (the oracle port on the same rows) with a lockstep `parity` block, clocks.

  c1  run_SRM.py:21-36      ERM / BCE / l1_reg 0.01 on scikit-learn synthetic 10000 x 1000 -> 6000 x 1000 train rows,
                            tol 1e-6 — the one configuration the reference itself runs end to end (126 iterations,
                            26-29 s here: tests/golden/c1_trajectory.npz); 48 MB, L2-resident: latency-bound
  c3  run_EHRM.py:27-39     EHRM (CPT spectra, B = -5) / BCE / l2_reg 0.01 on planted 4M x 500, rows sharded over N GPUs
  c4  run_AoRR_ratio.py:32-46  AoRR / hinge (or --loss binary_cross_entropy) / l2_reg 1e-4 on planted 2M x 200 + intercept
                            column, one solve per ratio pair args in {[.1,.9],[.2,.8],[.3,.7],[.4,.6]}
  c5  batched lambda grid   256 values of l1_reg log-spaced in [1e-4, 1], superquantile(0.8) / BCE on planted 100k x 1000,
                            32 instances per GPU (instances sharded over the GPUs, no communication: weak scaling)
"""
import contextlib
import io
import json
import os
import time

import numpy as np

UNIT = "iterations/s"

CONFIGS = {
    "c1": dict(title="ERM BCE l1_reg=0.01 ADMM on scikit-learn synthetic 6000 x 1000 fp64, tol 1e-6 (BASELINE configs[0], "
                     "run_SRM.py:21-36)",
               n=6000, d=1000, data="sklearn", wf="erm", sweeps=[None], loss="binary_cross_entropy", B=None,
               reg=dict(l1_reg=0.01), intercept=False, tol=1e-6, max_total=200, cpu_iters=None),
    "c3": dict(title="EHRM (CPT spectra, B=-5) BCE l2_reg=0.01 ADMM on planted 4M x 500 fp64 (BASELINE configs[2], "
                     "run_EHRM.py:27-39)",
               n=4_000_000, d=500, data="planted", wf="ehrm", sweeps=[None], loss="binary_cross_entropy", B=-5,
               reg=dict(l2_reg=0.01), intercept=False, tol=1e-6, max_total=120, cpu_iters=8),
    "c4": dict(title="AoRR ranked-range ADMM l2_reg=1e-4 on planted 2M x 200 (+ intercept column) fp64, ratio sweep "
                     "args in {[.1,.9],[.2,.8],[.3,.7],[.4,.6]} (BASELINE configs[3], run_AoRR_ratio.py:32-46)",
               n=2_000_000, d=200, data="planted", wf="aorr", sweeps=[[0.1, 0.9], [0.2, 0.8], [0.3, 0.7], [0.4, 0.6]],
               loss="hinge", B=None, reg=dict(l2_reg=1e-4), intercept=True, tol=1e-6, max_total=120, cpu_iters=8),
}


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)) / max(np.linalg.norm(b), 1e-300))


def _host_data(B, torch, cfg, dev, lo, hi):
    """(X, y) host arrays of rows [lo, hi) of the configuration's data set"""
    n, d = cfg["n"], cfg["d"]
    if cfg["data"] == "sklearn":
        from sklearn.model_selection import train_test_split
        from src.util.load_data import get_data

        X, y = get_data("synthetic", num_row=10000, num_feature=d, seed=17)
        Xtr, _, ytr, _ = train_test_split(X, y, test_size=0.4, random_state=17)
        assert Xtr.shape == (n, d)
        return np.ascontiguousarray(Xtr[lo:hi]), ytr[lo:hi].astype(np.float64)
    Xh, yh = B.gen_rows_device(torch, dev, lo, hi, n, d, pin=True)
    X, y = Xh.numpy(), yh.numpy().reshape(-1, 1)
    if cfg["intercept"]:  # run_AoRR_ratio.py:40-41
        X = np.hstack((X, np.ones((X.shape[0], 1))))
    return X, y


def run(args, B):
    """generic single-instance configurations (c1, c3, c4); B is the bench module (shared helpers)"""
    import torch
    import torch.distributed as dist

    from oracle import rbl_oracle as O  # cpu_baseline / parity leg only
    from rbl_b200 import _cabi
    from rbl_b200.engine import shard_bounds
    from src.optim.algorithms import ADMMmethod, Optimizer

    cfg = dict(CONFIGS[args.config])
    if args.n:
        cfg["n"] = args.n
    if args.loss:
        cfg["loss"] = args.loss
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, W, K = cfg["n"], args.warmup, args.steps
    lo, hi = shard_bounds(n, world, rank)
    X, y = _host_data(B, torch, cfg, dev, lo, hi)
    d_eff = X.shape[1]
    shard = dict(row_lo=lo, n_global=n) if world > 1 else {}
    quiet = contextlib.redirect_stdout(io.StringIO())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(sweep_args, Xa, ya, **extra):
        return ADMMmethod(Xa, ya, cfg["wf"], cfg["loss"], B=cfg["B"], args=sweep_args, max_iter=100_000, tol=cfg["tol"],
                          **cfg["reg"], **extra)

    # process warm-up (untimed): module load, lazy kernel loading — a tiny solve of the same kind and width
    rng_w = np.random.default_rng(5)
    Xw = rng_w.standard_normal((4096, d_eff))
    yw = np.sign(Xw[:, 0] + 0.1 * rng_w.standard_normal(4096)).reshape(-1, 1)
    with quiet:
        warm = make(cfg["sweeps"][0], Xw, yw)
        warm.advance(0, 6)
    warm.engine.close()
    del warm
    torch.cuda.synchronize()

    peak, peak_kind = B.measured_peak_gbs()
    sampler = B.ClockSampler(local_rank)
    sweeps_out, tot_steps_s, tot_e2e_s, tot_iters, tot_timed_iters, launches_timed = [], 0.0, 0.0, 0, 0, 0
    roofline = None
    if rank == 0:
        sampler.start()
    share = bool(getattr(args, "share_design", False)) and world == 1 and len(cfg["sweeps"]) > 1
    first = None
    for sweep_args in cfg["sweeps"]:
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        with quiet:
            if share and first is not None:   # --share-design: solves 2.. borrow the first solver's D, G and D^T
                s = make(sweep_args, None, None, _share=first)
            else:
                s = make(sweep_args, X, y, _shard=shard)
        ev[1].record()
        with quiet:
            it, done = s.advance(0, W)
        barrier()
        l0 = s.engine.launches
        ev[2].record()
        with quiet:
            if not done:
                it, done = s.advance(W, K)
        ev[3].record()
        barrier()
        launches_timed += s.engine.launches - l0
        with quiet:
            if not done:
                it, done = s.advance(W + K, max(0, cfg["max_total"] - (W + K)))
        ev[4].record()
        barrier()
        t_build = ev[0].elapsed_time(ev[1]) / 1e3
        t_steps = ev[2].elapsed_time(ev[3]) / 1e3
        t_all = ev[0].elapsed_time(ev[4]) / 1e3
        if world > 1:
            t = torch.tensor([t_build, t_steps, t_all], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_build, t_steps, t_all = (float(v) for v in t)
        timed = min(K, max(0, it - W))
        with quiet:
            obj = s.objective.get_arrogate_loss(torch.from_numpy(s.w).double())
        sweeps_out.append({"args": sweep_args, "iterations": it, "converged": bool(done),
                           "timed_iterations": timed, "ms_per_step": 1e3 * t_steps / max(timed, 1),
                           "upload_and_build_s": round(t_build, 4), "solve_incl_upload_s": round(t_all, 4),
                           "objective": obj, "primal": s.primal_feasibility, "dual": s.dual_feasibility,
                           "active_rows": dict(s.engine.active_stats), "dual_pass": dict(s.engine.dual_stats),
                           "ehrm_choice": dict(s.engine.ehrm_stats) if cfg["wf"] == "ehrm" else None,
                           "lbfgs_last": s.last_info if "l2_reg" in cfg["reg"] else None})
        tot_steps_s += t_steps
        tot_timed_iters += timed
        tot_e2e_s += t_all
        tot_iters += it
        if roofline is None:
            # the D-reading kernel of these configurations is the streaming pass (every row is active for CPT /
            # ERM spectra; the l2 dual pass is dense): timed alone, 20 launches between events on its stream
            eng = s.engine
            nl = hi - lo
            xd = eng.w.clone()

            def one():
                _cabi.check(eng.lib.rbl_fused_pass(eng.h, eng.D.data_ptr(), xd.data_ptr(), eng.b.data_ptr(),
                                                   eng.r.data_ptr(), eng.red.data_ptr(), eng._stream()))
            for _ in range(3):
                one()
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(20):
                one()
            p1.record()
            torch.cuda.synchronize()
            t_pass = p0.elapsed_time(p1) / 1e3 / 20
            alg = nl * d_eff * 8 + (2 * nl + 2 * d_eff) * 8
            z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            z0.record()
            for _ in range(5):
                eng.z_step(s.rho)
            z1.record()
            torch.cuda.synchronize()
            t_z = z0.elapsed_time(z1) / 1e3 / 5
            eng.Dw_valid = False
            roofline = {"bound": "hbm", "kernel": "rbl_pass_kernel (fused r = b - D x, ||r||^2, D^T r over all rows)",
                        "achieved": alg / t_pass / 1e9, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                        "frac": alg / t_pass / 1e9 / peak, "traffic": None,
                        "algorithmic_bytes_per_launch": alg, "launch_ms": 1e3 * t_pass,
                        "note": ("D is %.0f MB per GPU: " % (nl * d_eff * 8 / 1e6))
                                + ("L2-resident (126 MB L2), the step is latency-bound" if nl * d_eff * 8 < 100e6
                                   else "streams from HBM every pass"),
                        "zstep": {"ms": 1e3 * t_z, "keys_per_s": n / t_z, "algorithmic_bytes": 68 * n,
                                  "frac_of_hbm_peak": 68 * n / t_z / 1e9 / peak}}
        if share and first is None:
            first = s
            continue
        s.engine.close()
        del s
        if not share:
            torch.cuda.empty_cache()
    if first is not None:
        first.engine.close()
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- cpu_baseline + parity (N = 1): the oracle on the same rows in free-running lockstep with a GPU solver ----
    cpu = parity = None
    if world == 1 and not args.no_cpu:
        cores = B.use_all_host_threads()
        kc = cfg["cpu_iters"] or (W + K)
        wc = min(W, max(0, kc - 3))
        sweep_args = cfg["sweeps"][0]
        with quiet:
            par = make(sweep_args, X, y)
        o = O.OracleADMM(X, y, cfg["wf"], cfg["loss"], B=cfg["B"], args=sweep_args, max_iter=100_000, tol=cfg["tol"],
                         **cfg["reg"])
        import copy

        o1 = copy.copy(o)   # shares D / DTD; its state is overwritten with the device state before every step
        rows, dt = [], 0.0
        for i in range(kc):
            o1.w, o1.z, o1.lam, o1.rho = (par.w.reshape(-1).copy(), par.z.reshape(-1).copy(),
                                          par.lagrangian.reshape(-1).copy(), par.rho)
            t0 = time.perf_counter()
            o.step()
            if i >= wc:
                dt += time.perf_counter() - t0
            with quiet:
                Optimizer.main_loop(par, i, 0.0, False)
            o1.step()
            zg = par.engine.z.cpu().numpy()
            rows.append({"iteration": i, "rel_w": _rel(par.w, o.w), "rel_z": _rel(zg, o.z),
                         "rel_w_one_step": _rel(par.w, o1.w), "rel_z_one_step": _rel(zg, o1.z),
                         "rel_rho": abs(float(par.rho) - float(o.rho)) / float(o.rho)})
        with quiet:
            obj_gpu = par.objective.get_arrogate_loss(torch.from_numpy(par.w).double())
        obj_cpu = o.objective()
        v = (kc - wc) / dt
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle port (numpy/BLAS on all host threads, C stack-PAV, scipy L-BFGS-B / numpy FISTA fp64) "
                         f"on the SAME {n} x {d_eff} rows, ADMM iterations {wc}..{kc - 1} of the first solve of the "
                         f"configuration ({dt:.1f} s of CPU work), no scaling"}
        parity = {"tolerance": 1e-9, "iterations_compared": kc,
                  "max_rel_w_one_step": max(r["rel_w_one_step"] for r in rows),
                  "max_rel_z_one_step": max(r["rel_z_one_step"] for r in rows),
                  "max_rel_w_free_running": max(r["rel_w"] for r in rows),
                  "max_rel_z_free_running": max(r["rel_z"] for r in rows),
                  "max_rel_rho": max(r["rel_rho"] for r in rows),
                  "objective_gpu": obj_gpu, "objective_cpu": obj_cpu,
                  "rel_objective": abs(obj_gpu - obj_cpu) / abs(obj_cpu),
                  "per_iteration": [{k: (float("%.3e" % x) if isinstance(x, float) else x) for k, x in r.items()}
                                    for r in rows],
                  "how": "every iteration twice: (one step) the CPU oracle restarted from the device state, and (free "
                         "running) a second oracle that never sees the device state.  The l2 w-step is scipy's "
                         "L-BFGS-B stopped at its default gtol 1e-5 / ftol 2.2e-9 on both sides: rounding-level "
                         "differences in f and g move its stopping point by ~1e-12 per call, which the free-running "
                         "pair accumulates (both trajectories are equally valid); the one-step figure is the parity "
                         "of the device path itself"}
        parity["ok"] = bool(parity["max_rel_w_one_step"] <= 1e-9 and parity["max_rel_z_one_step"] <= 1e-9
                            and parity["rel_objective"] <= 1e-9)
        if cfg["loss"] == "hinge":
            parity["note"] = ("hinge: the oracle and the GPU use the closed-form prox; the reference's early-exit "
                              "bisection (individual_solver.py:15-42) is waived (SURVEY §8a iii)")
        par.engine.close()

    value = tot_timed_iters / tot_steps_s
    out = {"metric": "admm_iters_per_sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": cfg["title"], "name": args.config, "n": n, "d": d_eff, "loss": cfg["loss"],
                      "rows_per_gpu": hi - lo, "parallelism": f"rows sharded x{world}" if world > 1 else "single GPU",
                      "timed_iterations": f"{W}..{W + K - 1} of every solve from the reference's initial state",
                      "tolerance": cfg["tol"], "iteration_cap": cfg["max_total"],
                      "l2_flush": ("none: D (%.0f MB per GPU) is " % ((hi - lo) * d_eff * 8 / 1e6))
                                  + ("smaller than L2 — the reference's own problem size, reported as it is"
                                     if (hi - lo) * d_eff * 8 < 126e6 else "far larger than the 126 MB L2"),
                      "solves": sweeps_out},
           "e2e": {"value": tot_iters / tot_e2e_s, "unit": UNIT,
                   "h2d_bytes_per_step": (1 if share else len(cfg["sweeps"])) * (X.size + y.size) * 8 / max(tot_iters, 1),
                   "design_matrix": ("uploaded once, shared by the solves of the sweep (ADMMmethod(_share=first))"
                                     if share else "uploaded by every solve (one ADMMmethod(X, y, ...) per solve, "
                                                   "like the reference's drivers)"),
                   "d2h_bytes_per_step": (d_eff + 16) * 8,
                   "note": "ADMMmethod(X, y, ...) on host numpy arrays (upload, D = -y*X, G = D^T D) + the loop to the "
                           "stop test or the iteration cap, w and the residuals read back every iteration; all "
                           "iterations / (upload + build + solve), summed over the solves of the configuration"},
           "gpu_launches": launches_timed, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
    if args.config == "c1":
        try:
            g = np.load(os.path.join(B.ROOT, "tests", "golden", "c1_trajectory.npz"))
            out["reference_as_shipped"] = {
                "what": "the reference's own ADMMmethod (+ the get_opt shim, float64 FISTA) on this configuration, run "
                        "in the build container by oracle/gen_golden.py::config1",
                "iterations": int(g["iterations"]), "wall_s": float(g["ref_wall_s"]), "cores": int(g["ref_cores"]),
                "iterations_per_s": int(g["iterations"]) / float(g["ref_wall_s"]), "objective": float(g["objective"])}
        except Exception:  # noqa: BLE001
            pass
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args, B):
    """--impl reference --config c1|c3|c4: the oracle port on the host cores, same data, W warm-up + K timed steps"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    from oracle import rbl_oracle as O

    cfg = dict(CONFIGS[args.config])
    if args.n:
        cfg["n"] = args.n
    if args.loss:
        cfg["loss"] = args.loss
    cores = B.use_all_host_threads()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else None
    if dev is None and cfg["data"] != "sklearn":
        raise SystemExit("the planted data blocks are generated on the device: no CUDA device here")
    X, y = _host_data(B, torch, cfg, dev, 0, cfg["n"])
    W, K = args.warmup, args.steps
    o = O.OracleADMM(X, y, cfg["wf"], cfg["loss"], B=cfg["B"], args=cfg["sweeps"][0], max_iter=100_000, tol=cfg["tol"],
                     **cfg["reg"])
    dt, done_at = 0.0, None
    for i in range(W + K):
        t0 = time.perf_counter()
        fin = o.step()
        if i >= W:
            dt += time.perf_counter() - t0
        if fin:
            done_at = i + 1
            break
    timed = (done_at or (W + K)) - W
    v = timed / dt
    print(json.dumps({
        "impl": "reference", "metric": "admm_iters_per_sec", "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["title"], "name": args.config, "n": cfg["n"], "d": X.shape[1], "loss": cfg["loss"],
                   "same_data_as_gpu_arm": True, "timed_iterations": timed},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle port on the same {cfg['n']} x {X.shape[1]} rows, iterations {W}..{W + timed - 1} "
                                   f"({dt:.1f} s of CPU work)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ---- c5: batched lambda grid, instances sharded over the GPUs --------------------------------------------------------
def run_c5(args, B):
    import torch
    import torch.distributed as dist

    from oracle import rbl_oracle as O
    from rbl_b200.batched import BatchedADMM

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, W, K = args.n or 100_000, 1000, args.warmup, args.steps
    per_gpu = args.instances_per_gpu
    total = per_gpu * world
    # the 256-point grid of the configuration; with fewer than 8 GPUs the job runs the first 32 N points of it, spread
    # evenly over the whole range so that every N sees the same mix of regularisation strengths
    grid = np.logspace(-4, 0, 256)
    regs = grid[np.linspace(0, 255, total).round().astype(int)] if total < 256 else grid[:total]
    Xh, yh = B.gen_rows_device(torch, dev, 0, n, n, d, pin=True)   # every rank holds the same shared D
    X, y = Xh.numpy(), yh.numpy().reshape(-1, 1)
    quiet = contextlib.redirect_stdout(io.StringIO())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    b = BatchedADMM(X, y, "superquantile", "binary_cross_entropy", l1_regs=list(regs), args=[0.8], max_iter=10_000,
                    tol=1e-6, mode=args.batch_mode)
    ev[1].record()
    for _ in range(W):
        b.step()
    barrier()
    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lc0 = int(b.eng.lib.rbl_launch_count()) + sum(getattr(c, "_graph_replays", 0) * getattr(c, "_graph_launches", 0)
                                                  for c in getattr(b, "inst", []))
    act0 = b.n_active
    ev[2].record()
    inst_iters = 0
    for _ in range(K):
        inst_iters += b.n_active
        b.step()
    ev[3].record()
    barrier()
    lc1 = int(b.eng.lib.rbl_launch_count()) + sum(getattr(c, "_graph_replays", 0) * getattr(c, "_graph_launches", 0)
                                                  for c in getattr(b, "inst", []))
    clocks = sampler.stop() if rank == 0 else None
    t_build = ev[0].elapsed_time(ev[1]) / 1e3
    t_steps = ev[2].elapsed_time(ev[3]) / 1e3
    t_e2e = ev[0].elapsed_time(ev[3]) / 1e3
    tot = torch.tensor([float(inst_iters), float(inst_iters + W * act0)], dtype=torch.float64, device=dev)
    tm = torch.tensor([t_steps, t_e2e, t_build], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    inst_iters_all, inst_iters_e2e = float(tot[0]), float(tot[1])
    t_steps, t_e2e, t_build = (float(v) for v in tm)
    if rank != 0:
        b.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_kind = B.measured_peak_gbs()
    act = [c.active_stats["rows"] / max(1, c.active_stats["calls"]) for c in getattr(b, "inst", [])]
    cpu = parity = None
    if world == 1 and not args.no_cpu:
        cores = B.use_all_host_threads()
        # three instances of the grid (weakest, middle, strongest regularisation) in lockstep with the oracle
        picks = sorted({0, per_gpu // 2, per_gpu - 1})
        bb = BatchedADMM(X, y, "superquantile", "binary_cross_entropy", l1_regs=[float(regs[j]) for j in picks],
                         args=[0.8], max_iter=10_000, tol=1e-6, mode=args.batch_mode)
        os_ = [O.OracleADMM(X, y, "superquantile", "binary_cross_entropy", l1_reg=float(regs[j]), args=[0.8],
                            max_iter=10_000, tol=1e-6) for j in picks]
        kc, dt, worst_w, worst_z = min(W + K, 10), 0.0, 0.0, 0.0
        per_inst = [{"l1_reg": float(regs[j]), "max_rel_w": 0.0, "max_rel_z": 0.0, "max_abs_z": 0.0} for j in picks]
        for i in range(kc):
            bb.step()
            for k, o in enumerate(os_):
                t0 = time.perf_counter()
                o.step()
                dt += time.perf_counter() - t0
                w, z, _, _ = bb.state(k)
                ew = _rel(w, o.w) if np.linalg.norm(o.w) > 0 else float(np.linalg.norm(w))
                ez = _rel(z, o.z)
                pi = per_inst[k]
                pi["max_rel_w"], pi["max_rel_z"] = max(pi["max_rel_w"], ew), max(pi["max_rel_z"], ez)
                pi["max_abs_z"] = max(pi["max_abs_z"], float(np.max(np.abs(z - o.z))))
                pi["nnz_w"] = int(np.count_nonzero(o.w))
                worst_w, worst_z = max(worst_w, ew), max(worst_z, ez)
        v = kc * len(os_) / dt
        cpu = {"value": v, "unit": "instance-iterations/s", "cores": cores, "kind": "port",
               "sample": f"oracle port, one instance at a time, on the same {n} x {d} rows: {len(os_)} instances of the "
                         f"grid (l1_reg {[float('%.3g' % regs[j]) for j in picks]}) x iterations 0..{kc - 1} "
                         f"({dt:.1f} s of CPU work)"}
        parity = {"tolerance": 1e-9, "instances_compared": len(os_), "iterations_compared": kc,
                  "max_rel_w": worst_w, "max_rel_z": worst_z,
                  # an instance whose regulariser kills w has z ~ 1e-7 in norm: its z is judged on the absolute error
                  "ok": bool(all(pi["max_rel_w"] <= 1e-9 and (pi["max_rel_z"] <= 1e-9 or pi["max_abs_z"] <= 1e-12)
                                 for pi in per_inst)),
                  "per_instance": per_inst,
                  "how": "free-running lockstep of batched GPU instances and per-instance CPU oracles"}
        bb.close()
    value = inst_iters_all / t_steps
    out = {"metric": "admm_instance_iters_per_sec", "value": value, "unit": "instance-iterations/s", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": 1e3 * t_steps / K, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"batched lambda-grid SRM superquantile(0.8) BCE, l1_reg log-spaced in [1e-4, 1], shared "
                                  f"planted {n} x {d} fp64 design, {per_gpu} instances per GPU x {world} GPU(s) = {total} "
                                  f"of the 256 (BASELINE configs[4])",
                      "name": "c5", "instances_per_gpu": per_gpu, "instances_total": total, "batch_mode": b.mode,
                      "parallelism": f"instances sharded x{world}, no communication",
                      "step": "one ADMM iteration of every still-active instance of the rank",
                      "instances_still_active_after_timed_steps": int(b.n_active),
                      "mean_active_rows_per_instance": float(np.mean(act)) if act else None,
                      "l2_flush": "none: D (%.0f MB) exceeds L2 only together with the per-instance state; every "
                                  "instance re-reads its active rows of D" % (n * d * 8 / 1e6)},
           "e2e": {"value": inst_iters_e2e / t_e2e, "unit": "instance-iterations/s",
                   "h2d_bytes_per_step": (X.size + y.size) * 8 / (W + K), "d2h_bytes_per_step": per_gpu * 32,
                   "note": "BatchedADMM(X, y, ...) from host arrays (upload, D, G) + W + K batched iterations",
                   "upload_and_build_s": t_build},
           "gpu_launches": lc1 - lc0, "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": "rbl_gather_kernel per instance (active rows of the shared D)",
                        "achieved": None, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": None,
                        "traffic": None,
                        "note": "bytes per instance-iteration ~ active rows x d x 8; see the c2 line for the kernel's "
                                "own roofline (1.0 of the measured peak)",
                        "achieved_from_active_rows": (float(np.mean(act)) * d * 8 * value / world / 1e9) if act else None},
           "cpu_baseline": cpu, "parity": parity}
    if out["roofline"]["achieved_from_active_rows"] is not None:
        out["roofline"]["achieved"] = out["roofline"]["achieved_from_active_rows"]
        out["roofline"]["frac"] = out["roofline"]["achieved"] / peak
    print(json.dumps(out), flush=True)
    b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
