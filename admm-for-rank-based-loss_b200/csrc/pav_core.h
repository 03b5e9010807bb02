// pav_core.h — the merge step of the parallel pool-adjacent-violators (isotonic) prox.
//
// Replaces (semantics): src/util/pav.py:93-178 (PAV_solver.get_opt) and
// src/util/PAV_cpt.py:234-293 (PAV_solver_CPT.get_opt).  The reference sweeps: every pass merges
// each run of adjacent violators and re-solves; superquantile / AoRR spectra need Theta(n) passes
// because the pooled block swallows one sigma=0 neighbour per pass.  Here the same fixed point
// (the unique isotonic prox) is reached by a balanced tree of merges of depth log2(n):
//
//   level 0 : every element is a solved range, val[i] = prox(sigma_i, m_i)
//   level l : adjacent solved ranges L=[a,b), R=[b,c) are merged.  Inside a solved range val is
//             non-decreasing.  If val[b-1] <= val[b] nothing changes.  Otherwise exactly one new
//             pooled block [lo*, hi*) appears across the boundary; everything else is unchanged
//             ("clipping" property of isotonic merges):  z_i = min(z^L_i, t*) on L and
//             max(z^R_i, t*) on R, where t* minimises
//                 F(t) = sum_{i in L} f_i(min(z^L_i, t)) + sum_{i in R} f_i(max(z^R_i, t)),
//             f_i(z) = sigma_i*loss(z) + rho/2 (z - m_i)^2.   F is convex and
//                 F'(t) = Phi(t) = sum_{i in P(t)} f_i'(t),  P(t) = {L: val > t} u {R: val < t},
//             a monotone function whose breakpoints are the block values.  Because a block's sum of
//             f_i' depends only on (sum sigma, sum m, count), Phi at any breakpoint costs O(1) given
//             prefix sums of sigma and m plus two bound searches in val — so lo*, hi* are found by
//             binary search (O(log^2) reads), not by absorbing blocks one at a time.  That keeps the
//             worst case (one giant block: iteration 0 of every superquantile run) as cheap as the
//             best case (all singletons: ERM, or any converged run).
//
// Ties (hinge produces many blocks sitting exactly on the kink z = -1) are handled with one-sided
// limits: an L element joins the pooled block iff val >= t*  <=>  Phi(u+) >= 0, an R element iff
// val <= t*  <=>  Phi(u-) <= 0.  Pooling tied blocks never changes the pooled value.
//
// This header is pure scalar code (host + device) so tests/native/ can run the identical logic on
// the CPU against the oracle's sequential stack PAV.
#pragma once
#include "prox_core.h"

// exclusive prefix sums stored as double-double with n+1 entries
struct PrefixFlat {
    const double* hi;
    const double* lo;
    RBL_HDM dd_t get(int64_t i) const {
        dd_t r;
        r.hi = hi[i];
        r.lo = lo[i];
        return r;
    }
};

// chunk-local exclusive prefixes (n+1 entries; entry n closes a partial last chunk) + per-chunk
// offsets (nchunks+1 entries)
struct PrefixChunked {
    const double* loc_hi;
    const double* loc_lo;
    const double* off_hi;
    const double* off_lo;
    int shift;  // log2(chunk)
    RBL_HDM dd_t get(int64_t i) const {
        int64_t c = i >> shift;
        dd_t r;
        r.hi = off_hi[c];
        r.lo = off_lo[c];
        if (i & ((int64_t(1) << shift) - 1)) {
            dd_t l;
            l.hi = loc_hi[i];
            l.lo = loc_lo[i];
            r = dd_add(r, l);
        }
        return r;
    }
};

// first i in [lo,hi) with val[i] >= u  (hi if none)
RBL_HD int64_t pav_lower_bound(const double* val, int64_t lo, int64_t hi, double u) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (val[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// first i in [lo,hi) with val[i] > u  (hi if none)
RBL_HD int64_t pav_upper_bound(const double* val, int64_t lo, int64_t hi, double u) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (val[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// sum_{i in [l,r)} f_i'(u) with the one-sided loss derivative
template <class PS, class PM>
RBL_HD double pav_phi(int loss, double rho, const PS& ps, const PM& pm, int64_t l, int64_t r, double u, int side) {
    if (r <= l) return 0.0;
    double ssig = dd_diff(ps.get(r), ps.get(l));
    double sm = dd_diff(pm.get(r), pm.get(l));
    return ssig * rbl_loss_deriv(loss, u, side) + rho * ((double)(r - l) * u - sm);
}

template <class PS, class PM>
RBL_HD double pav_block_value(int loss, double rho, const PS& ps, const PM& pm, int64_t l, int64_t r) {
    double cnt = (double)(r - l);
    double ssig = dd_diff(ps.get(r), ps.get(l));
    double sm = dd_diff(pm.get(r), pm.get(l));
    return rbl_block_prox(loss, ssig / cnt, sm / cnt, rho);
}

// end of the run of values <= u that contains position p (val[p] <= u): first i in (p, limit) with
// val[i] > u, found by galloping so that the usual short run costs one or two reads
RBL_HD int64_t pav_run_end(const double* val, int64_t p, int64_t limit, double u) {
    int64_t prev = p, step = 1, probe = p + 1;
    while (probe < limit && val[probe] <= u) {
        prev = probe;
        step <<= 1;
        probe = prev + step;
    }
    if (probe > limit) probe = limit;
    return pav_upper_bound(val, prev + 1, probe, u);
}

// start of the run of values >= u that contains position p (val[p] >= u): first i in [limit, p] with
// val[i] >= u, galloping backwards
RBL_HD int64_t pav_run_start(const double* val, int64_t p, int64_t limit, double u) {
    int64_t prev = p, step = 1, probe = p - 1;
    while (probe >= limit && val[probe] >= u) {
        prev = probe;
        step <<= 1;
        probe = prev - step;
    }
    if (probe < limit - 1) probe = limit - 1;
    return pav_lower_bound(val, probe + 1, prev, u);
}

// Merge solved ranges [a,b) and [b,c).  Returns false if there is no violation at the boundary;
// otherwise the pooled block [*lo_out, *hi_out) and its value.
//
// Cost: the bound searches on the probe's own side gallop from the probe (runs are short), and the
// searches on the opposite side are monotone in the probe position, so their window [xlo, xhi]
// shrinks together with the outer binary search: O(log n) dependent reads per side in total.
template <class PS, class PM>
RBL_HD bool pav_merge_search(int loss, double rho, const double* val, const PS& ps, const PM& pm, int64_t a,
                             int64_t b, int64_t c, int64_t* lo_out, int64_t* hi_out, double* v_out) {
    if (!(val[b - 1] > val[b])) return false;
    // ---- left side: first p in [a,b) with Phi(val[p]+) >= 0; val[b-1] always qualifies
    int64_t lo = a, hi = b - 1;
    int64_t xlo = b, xhi = c;  // window of upper_bound(val[b..c), val[p]) for the remaining probes
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        double u = val[mid];
        int64_t l = pav_run_end(val, mid, b, u);          // {L: val > u} = [l, b)
        int64_t r = pav_upper_bound(val, xlo, xhi, u);    // {R: val <= u} = [b, r)
        if (pav_phi(loss, rho, ps, pm, l, r, u, +1) >= 0.0) {
            hi = mid;
            xhi = r;
        } else {
            lo = mid + 1;
            xlo = r;
        }
    }
    int64_t lo_star = lo;
    // ---- right side: first p in (b,c) with Phi(val[p]-) > 0; val[b] never qualifies
    lo = b + 1;
    hi = c;
    xlo = a;
    xhi = b;  // window of lower_bound(val[a..b), val[p])
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        double u = val[mid];
        int64_t l = pav_lower_bound(val, xlo, xhi, u);    // {L: val >= u} = [l, b)
        int64_t r = pav_run_start(val, mid, b, u);        // {R: val < u} = [b, r)
        if (pav_phi(loss, rho, ps, pm, l, r, u, -1) > 0.0) {
            hi = mid;
            xhi = l;
        } else {
            lo = mid + 1;
            xlo = l;
        }
    }
    int64_t hi_star = lo;
    // snap to whole runs of equal values (probes inside a run give the same answer, but rounding
    // in Phi near a tie may split one; equal values pool for free)
    lo_star = pav_run_start(val, lo_star, a, val[lo_star]);
    hi_star = pav_run_end(val, hi_star - 1, c, val[hi_star - 1]);
    double v = pav_block_value(loss, rho, ps, pm, lo_star, hi_star);
    // exact arithmetic guarantees val[lo*-1] <= v <= val[hi*]; clamp so rounding can never break the
    // "non-decreasing inside a solved range" invariant the bound searches rely on
    if (lo_star > a && v < val[lo_star - 1]) v = val[lo_star - 1];
    if (hi_star < c && v > val[hi_star]) v = val[hi_star];
    *lo_out = lo_star;
    *hi_out = hi_star;
    *v_out = v;
    return true;
}
