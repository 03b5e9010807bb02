"""Python restatement of the bucket search of ss_partition_ranked_kernel (csrc/sort_kernels.cu): the bucket of a key is
guessed from the row's previous rank, verified against the sorted splitters, and found by a gallop + bisection from the
guess when the row has moved.  Whatever the guess, the answer must be the largest j with spl[j] <= key (spl[0] = 0),
which is what the row-order kernel's full bisection returns.  Run by tests/test_host.py."""
import numpy as np


def guess_bucket(i, n, nb):
    g = ((i + 1) * nb - 1) // n            # the largest j with floor(j n / nb) <= i
    return min(max(g, 0), nb - 1)


def bucket_from_guess(spl, nb, k, g):
    if spl[g] <= k:
        lo, hi, step = g, g + 1, 1
        while hi < nb and spl[hi] <= k:
            lo = hi
            hi = min(nb, lo + step)
            step <<= 1
    else:                                   # spl[0] = 0 <= every key
        hi, lo, step = g, g - 1, 1
        while lo > 0 and spl[lo] > k:
            hi = lo
            lo = max(0, hi - step)
            step <<= 1
    while hi - lo > 1:
        mid = (lo + hi) >> 1
        if spl[mid] <= k:
            lo = mid
        else:
            hi = mid
    return lo


def run_cases():
    rng = np.random.default_rng(3)
    for nb in (16, 64, 1024, 4096):
        for kind in ("distinct", "duplicates", "few"):
            if kind == "distinct":
                spl = np.sort(rng.choice(1 << 40, size=nb, replace=False)).astype(np.uint64)
            elif kind == "duplicates":
                spl = np.sort(rng.integers(0, nb // 4 + 2, size=nb)).astype(np.uint64) * np.uint64(1000)
            else:
                spl = np.sort(rng.integers(0, 3, size=nb)).astype(np.uint64) * np.uint64(7)
            spl[0] = 0
            keys = np.concatenate([rng.integers(0, int(spl[-1]) + 50, size=400).astype(np.uint64),
                                   spl[rng.integers(0, nb, size=100)],            # keys equal to splitters
                                   np.array([0, int(spl[-1]), (1 << 64) - 1], dtype=np.uint64)])
            ref = np.searchsorted(spl, keys, side="right") - 1
            for k, r in zip(keys, ref):
                for g in (int(r), max(int(r) - 1, 0), min(int(r) + 1, nb - 1), 0, nb - 1, int(rng.integers(0, nb))):
                    assert bucket_from_guess(spl, nb, k, g) == r, (nb, kind, int(k), g, int(r))
    # the guess itself: rank i of n lies between the ranks floor(j n / nb) the splitters were taken from
    for n, nb in ((3000, 16), (70001, 128), (1_000_000, 1024), (1 << 22, 4096)):
        pos = (np.arange(nb, dtype=np.int64) * n) // nb
        for i in np.concatenate([pos, pos - 1, pos + 1, [0, n - 1], rng.integers(0, n, size=200)]):
            if 0 <= i < n:
                assert guess_bucket(int(i), n, nb) == np.searchsorted(pos, i, side="right") - 1, (n, nb, int(i))


if __name__ == "__main__":
    run_cases()
    print("ok")
