"""world_size-2 gloo worker for tests/test_host.py::test_row_sharded_collectives_on_gloo.

Exercises the SAME sharding arithmetic and collective sequence the GPU engine uses (engine.shard_bounds,
all_gather_into_tensor with padding for uneven shards, all_reduce of the d+2 reduce buffer), with the
oracle standing in for the per-shard kernels (test infrastructure; no product code computes on CPU)."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from oracle import rbl_oracle as O
from rbl_b200.engine import shard_bounds


def gather_rows(local, n_global, world):
    base, extra = divmod(n_global, world)
    out = torch.empty(n_global, dtype=local.dtype)
    if extra == 0:
        dist.all_gather_into_tensor(out, local)
    else:
        pad = base + 1
        buf = torch.zeros(world * pad, dtype=local.dtype)
        mine = torch.zeros(pad, dtype=local.dtype)
        mine[: local.numel()] = local
        dist.all_gather_into_tensor(buf, mine)
        for r in range(world):
            lo, hi = shard_bounds(n_global, world, r)
            out[lo:hi] = buf[r * pad: r * pad + (hi - lo)]
    return out


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same data on every rank; each keeps its rows
    n, d = 1001, 12             # uneven shards on purpose
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    D = -y[:, None] * X
    w = rng.normal(size=d) * 0.2
    lam = rng.normal(size=n) * 1e-3
    rho = 1e-3
    sig = O.spectrum("superquantile", n, [0.8])
    lo, hi = shard_bounds(n, world, rank)
    Dl, laml = D[lo:hi], lam[lo:hi]
    # z-step: local margins -> all-gather -> replicated sort + PAV -> local scatter
    m_local = torch.from_numpy(Dl @ w - laml / rho)
    m = gather_rows(m_local, n, world).numpy()
    perm = np.argsort(m, kind="stable")
    zs = O.pav_prox("binary_cross_entropy", sig, m[perm], rho)
    z_local = np.zeros(hi - lo)
    mask = (perm >= lo) & (perm < hi)
    z_local[perm[mask] - lo] = zs[mask]
    z_ref = O.z_step(D, w, lam, rho, sig, "binary_cross_entropy")
    assert np.array_equal(m, D @ w - lam / rho) or np.allclose(m, D @ w - lam / rho, rtol=0, atol=1e-15)
    assert np.max(np.abs(z_local - z_ref[lo:hi])) < 1e-13
    # fused pass partials -> all-reduce
    b = z_ref + lam / rho
    r_local = b[lo:hi] - Dl @ w
    red = torch.from_numpy(np.concatenate([Dl.T @ r_local, [r_local @ r_local, 0.0]]))
    dist.all_reduce(red)
    r = b - D @ w
    assert np.allclose(red[:d].numpy(), D.T @ r, rtol=1e-12, atol=1e-12)
    assert abs(float(red[d]) - r @ r) < 1e-10 * (r @ r)
    # every rank sees bit-identical reduced values => identical branch decisions in the FISTA state machine
    chk = [torch.zeros_like(red) for _ in range(world)]
    dist.all_gather(chk, red)
    assert all(torch.equal(chk[0], c) for c in chk)
    # ---- Gram-mode data flow (engine.gram / _pass_at / dual pass): G = sum of the per-rank D_l^T D_l (one
    # all-reduce at construction); g0 = D^T (z - m) from each rank's ACTIVE rows only (z != m), all-reduced;
    # the FISTA trial identities on G; the primal residual as an all-reduced scalar
    G = torch.from_numpy(Dl.T @ Dl)
    dist.all_reduce(G)
    assert np.allclose(G.numpy(), D.T @ D, rtol=1e-13, atol=1e-12)
    delta_l = z_local - m[lo:hi]
    act = np.flatnonzero(delta_l != 0.0)
    assert 0 < len(act) < (hi - lo)                       # superquantile: most rows are untouched by the prox
    g0 = torch.from_numpy(np.concatenate([Dl[act].T @ delta_l[act], [delta_l[act] @ delta_l[act]]]))
    dist.all_reduce(g0)
    assert np.allclose(g0[:d].numpy(), D.T @ r, rtol=1e-11, atol=1e-12)        # b - D w == z - m
    assert abs(float(g0[d]) - r @ r) < 1e-10 * (r @ r)
    beta = w + 1e-2 * rng.normal(size=d)                  # same on every rank (same generator state)
    g_beta = g0[:d].numpy() - G.numpy() @ (beta - w)      # D^T (b - D beta) = g0 - G (beta - w)
    assert np.allclose(g_beta, D.T @ (b - D @ beta), rtol=1e-10, atol=1e-11)
    dl = beta - w
    ss = float(g0[d]) - 2 * dl @ g0[:d].numpy() + dl @ (G.numpy() @ dl)
    assert abs(ss - np.sum((b - D @ beta) ** 2)) < 1e-10 * ss
    prim = torch.tensor([float(np.sum((z_local - Dl @ beta) ** 2))], dtype=torch.float64)
    dist.all_reduce(prim)
    assert abs(float(prim[0]) - np.sum((z_ref - D @ beta) ** 2)) < 1e-11 * float(prim[0])
    # ---- test-set metrics of a row-sharded test set (rbl_b200.metrics.DeviceTestSet(sharded=True)): the 16 numbers
    # are sums over rows -> one all-reduce; the host formulas (product code) then give the whole-set statistics
    from rbl_b200.metrics import statistics_from_counts
    grp = (rng.random(n) < 0.4).astype(int)
    wm = rng.normal(size=d) * 0.5
    cnt, sb, sbl = O.confusion_by_group(wm, X[lo:hi], y[lo:hi], grp[lo:hi], 0.45)
    c16 = torch.zeros(16, dtype=torch.float64)
    c16[0] = O.calculate_accuracy(wm, X[lo:hi], y[lo:hi], 0.45) * (hi - lo)
    c16[1], c16[14], c16[15] = hi - lo, sb, sbl
    c16[2:14] = torch.from_numpy(cnt.reshape(-1))
    dist.all_reduce(c16)
    assert c16[1] == n and abs(float(c16[0]) / n - O.calculate_accuracy(wm, X, y, 0.45)) < 1e-15
    assert np.allclose(statistics_from_counts(c16.numpy()), O.calculate_statistics(wm, X, y, grp, 0.45),
                       rtol=1e-12, atol=1e-13)
    dist.barrier()
    print("OK", rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
