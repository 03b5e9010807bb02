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
    dist.barrier()
    print("OK", rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
