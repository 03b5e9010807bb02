"""Config-5 style measurement (BASELINE configs[4]): lambda-grid SRM solves on 100k x 1000 sharing one D,
instances sharded across GPUs with no communication.  Prints one JSON line (rank 0).

    python scripts/bench_batched.py [--instances 32] [--iters 20]
    python -m torch.distributed.run --nproc-per-node N scripts/bench_batched.py --instances 256
"""
import argparse, ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from rbl_b200.batched import BatchedADMM


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--d", type=int, default=1000)
    ap.add_argument("--mode", default="gram", choices=["gram", "stream"])
    ap.add_argument("--streams", type=int, default=8)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev); g.manual_seed(17)
    X = torch.randn(a.n, a.d, generator=g, dtype=torch.float64, device=dev)
    ws = torch.zeros(a.d, dtype=torch.float64, device=dev)
    ws[:10] = torch.randn(10, generator=g, dtype=torch.float64, device=dev)
    y = torch.sign(X @ ws + 0.1 * torch.randn(a.n, generator=g, dtype=torch.float64, device=dev)); y[y == 0] = 1
    regs = np.logspace(-4, 0, a.instances)
    b = BatchedADMM(X, y.reshape(-1, 1), "superquantile", "binary_cross_entropy", l1_regs=list(regs), args=[0.8],
                    max_iter=10_000, tol=1e-6, mode=a.mode, z_streams=a.streams)
    del X
    e = b.eng
    for _ in range(3):
        b.step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    p0 = b.fista_passes
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.iters):
        b.step()
    t1.record(); torch.cuda.synchronize()
    dt = t0.elapsed_time(t1) / 1e3
    passes = b.fista_passes - p0
    B = b.B
    if a.mode == "gram":
        out = {"workload": f"batched lambda-grid SRM superquantile(0.8) BCE l1 in [1e-4,1], n={a.n} d={a.d}, "
                           f"{a.instances} instances over {world} GPU(s) (BASELINE configs[4])",
               "mode": "gram: one child engine per instance, iteration graphs replayed concurrently on "
                       f"{len(b.streams)} streams", "instances_per_gpu": B, "admm_iterations_timed": a.iters,
               "instance_iterations_per_s_per_gpu": B * a.iters / dt, "ms_per_batched_iteration": 1e3 * dt / a.iters,
               "graphs": sum(1 for c in b.inst if c._graph is not None),
               "active_rows_mean": float(np.mean([c.active_stats["rows"] / max(1, c.active_stats["calls"])
                                                  for c in b.inst])),
               "nnz_last": [int(c.dual_stats["nnz_last"]) for c in b.inst][:8]}
        if world > 1:
            v = torch.tensor([out["instance_iterations_per_s_per_gpu"]], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(v)
            out["instance_iterations_per_s_total"] = float(v[0])
        if rank == 0:
            print(json.dumps(out))
        b.close()
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    # the multi-RHS pass alone: 8 instances per launch
    lams = (ctypes.c_double * B)(*([1.0] * B)); flags = (ctypes.c_int32 * B)(*([0] * B))
    _cabi.check(e.lib.rbl_fista_batch_begin(e.h, B, b.W.data_ptr(), lams, flags, 17.0, 0.0, 100000, e._stream()))
    _cabi.check(e.lib.rbl_fista_batch_steps(e.h, B, e.D.data_ptr(), b.Bv.data_ptr(), 2, e._stream()))
    torch.cuda.synchronize()
    K = 10
    q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
    q0.record()
    _cabi.check(e.lib.rbl_fista_batch_steps(e.h, B, e.D.data_ptr(), b.Bv.data_ptr(), K, e._stream()))
    q1.record(); torch.cuda.synchronize()
    groups = (B + 7) // 8
    t_pass = q0.elapsed_time(q1) / 1e3 / (K * groups)      # one multi-RHS step (pass + reduce + update + combine)
    flops = 2 * 2 * a.n * a.d * 8                            # two GEMMs, 8 right-hand sides
    out = {"workload": f"batched lambda-grid SRM superquantile(0.8) BCE l1 in [1e-4,1], n={a.n} d={a.d}, "
                       f"{a.instances} instances over {world} GPU(s) (BASELINE configs[4])",
           "instances_per_gpu": B, "admm_iterations_timed": a.iters,
           "instance_iterations_per_s_per_gpu": B * a.iters / dt, "ms_per_batched_iteration": 1e3 * dt / a.iters,
           "multi_rhs_passes_per_iteration": passes / a.iters,
           "multi_rhs_step_us": 1e6 * t_pass, "D_GBps_per_step": a.n * a.d * 8 / t_pass / 1e9,
           "fp64_TFLOPs_per_step": flops / t_pass / 1e12,
           "equivalent_single_instance_pass_us": 1e6 * t_pass / 8}
    if world > 1:
        v = torch.tensor([out["instance_iterations_per_s_per_gpu"]], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(v)
        out["instance_iterations_per_s_total"] = float(v[0])
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
