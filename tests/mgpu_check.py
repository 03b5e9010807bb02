"""Row-sharded parity check on N GPUs (launched with torchrun; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/mgpu_check.py

Every rank holds its contiguous rows; rank 0 compares the sharded solve with the unsharded CPU oracle."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)

from oracle import rbl_oracle as O  # noqa: E402
from rbl_b200.engine import shard_bounds  # noqa: E402
from src.optim.algorithms import ADMMmethod, Optimizer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    rng = np.random.default_rng(5)
    ok = True
    for (n, d, wf, args, loss, B, kw) in [
        (5003, 64, "superquantile", [0.8], "binary_cross_entropy", None, dict(l1_reg=0.01)),
        (4000, 201, "aorr", [0.2, 0.8], "hinge", None, dict(l2_reg=1e-4)),
        (3001, 50, "ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01)),
    ]:
        X = rng.normal(size=(n, d))
        ws = np.zeros(d)
        ws[:5] = rng.normal(size=5)
        y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
        lo, hi = shard_bounds(n, world, rank)
        s = ADMMmethod(X[lo:hi], y[lo:hi], wf, loss, B=B, args=args, max_iter=25, tol=1e-7,
                       _shard=dict(row_lo=lo, n_global=n), **kw)
        o = O.OracleADMM(X, y, wf, loss, B=B, args=args, max_iter=25, tol=1e-7, small_lasso=False, **kw)
        worst = 0.0
        for i in range(25):
            with contextlib.redirect_stdout(io.StringIO()):
                Optimizer.main_loop(s, i, 0.0, False)
            o.step()
            ew = np.linalg.norm(s.w.reshape(-1) - o.w) / np.linalg.norm(o.w)
            ez = np.linalg.norm(s.z.reshape(-1) - o.z[lo:hi]) / np.linalg.norm(o.z[lo:hi])
            worst = max(worst, ew, ez)
        obj = s.objective.get_arrogate_loss(torch.from_numpy(s.w).double())
        good = worst < 1e-9 and abs(obj - o.objective()) < 1e-9 * abs(obj)
        ok = ok and good
        if rank == 0:
            print(f"[{world} GPUs] {wf}/{loss}/{kw} n={n} d={d}: worst rel err {worst:.2e}, objective {obj:.12f} "
                  f"vs oracle {o.objective():.12f} -> {'OK' if good else 'FAIL'}", flush=True)
        if loss == "binary_cross_entropy":
            # test-set metrics of the row-sharded set: per-rank pass + one all-reduce of the 16 numbers
            from rbl_b200.metrics import DeviceTestSet
            grp = (np.random.default_rng(n).random(n) < 0.4).astype(int)
            ts = DeviceTestSet(X[lo:hi], y[lo:hi], group=grp[lo:hi], sharded=True)
            wv = s.w.reshape(-1)
            acc, st = ts.accuracy(wv), ts.statistics(wv)
            good_m = (acc == O.calculate_accuracy(wv, X, y) and
                      np.allclose(st, O.calculate_statistics(wv, X, y, grp), rtol=1e-12, atol=1e-13, equal_nan=True))
            ok = ok and good_m
            if rank == 0:
                print(f"[{world} GPUs] sharded test-set metrics n={n}: accuracy {acc:.6f} -> {'OK' if good_m else 'FAIL'}",
                      flush=True)
        s.engine.close()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
