// pav_core.h — the merge step of the parallel pool-adjacent-violators (isotonic) prox.
//
// Replaces (semantics): src/util/pav.py:93-178 (PAV_solver.get_opt) and
// src/util/PAV_cpt.py:234-293 (PAV_solver_CPT.get_opt).  The reference sweeps: every pass merges
// each run of adjacent violators and re-solves; superquantile / AoRR spectra need Theta(n) passes
// because the pooled block swallows one sigma=0 neighbour per pass.  Here the same fixed point
// (the unique isotonic prox) is reached by a balanced tree of merges of depth log2(n):
//
//   level 0 : every element is a solved range, val(i) = prox(sigma_i, m_i)
//   level l : adjacent solved ranges L=[a,b), R=[b,c) are merged.  Inside a solved range val is
//             non-decreasing.  If val(b-1) <= val(b) nothing changes.  Otherwise exactly one new
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

// read accessors for the block-value array: plain (host emulation, shared memory) or L1-bypassing
// (device global memory that other CTAs of the same launch have written)
struct ValPlain {
    const double* p;
    RBL_HDM double operator()(int64_t i) const { return p[i]; }
};
#if defined(__CUDACC__)
struct ValCG {
    const double* p;
    __device__ __forceinline__ double operator()(int64_t i) const { return __ldcg(p + i); }
};
#endif

// first i in [lo,hi) with val(i) >= u  (hi if none)
template <class V>
RBL_HD int64_t pav_lower_bound(const V& val, int64_t lo, int64_t hi, double u) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (val(mid) < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// first i in [lo,hi) with val(i) > u  (hi if none)
template <class V>
RBL_HD int64_t pav_upper_bound(const V& val, int64_t lo, int64_t hi, double u) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (val(mid) <= u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// upper_bound over [lo,hi) galloping up from lo (cheap when the answer is near lo)
template <class V>
RBL_HD int64_t pav_gallop_upper_from_lo(const V& val, int64_t lo, int64_t hi, double u) {
    int64_t step = 1, prev = lo, probe = lo;  // invariant: everything before prev is <= u
    while (probe < hi && val(probe) <= u) {
        prev = probe + 1;
        probe = lo + step;
        step <<= 1;
    }
    if (probe > hi) probe = hi;
    return pav_upper_bound(val, prev, probe, u);
}

// lower_bound over [lo,hi) galloping down from hi (cheap when the answer is near hi)
template <class V>
RBL_HD int64_t pav_gallop_lower_from_hi(const V& val, int64_t lo, int64_t hi, double u) {
    int64_t step = 1, prev = hi, probe = hi - 1;  // invariant: everything from prev on is >= u
    while (probe >= lo && val(probe) >= u) {
        prev = probe;
        probe = hi - 1 - step;
        step <<= 1;
    }
    if (probe < lo - 1) probe = lo - 1;
    return pav_lower_bound(val, probe + 1, prev, u);
}

// upper_bound over [lo,hi) started at an arbitrary position h (a guess of the answer): gallops away from h in the
// direction the value at h dictates — O(log |answer - h|) reads
template <class V>
RBL_HD int64_t pav_upper_bound_near(const V& val, int64_t lo, int64_t hi, double u, int64_t h) {
    if (hi <= lo) return lo;
    if (h < lo) h = lo;
    if (h > hi - 1) h = hi - 1;
    if (val(h) <= u) return pav_gallop_upper_from_lo(val, h, hi, u);  // everything before h is <= u as well
    int64_t step = 1, prev = h, probe = h - 1;                        // invariant: val(prev) > u
    while (probe >= lo && val(probe) > u) {
        prev = probe;
        step <<= 1;
        probe = h - step;
    }
    if (probe < lo) probe = lo - 1;
    return pav_upper_bound(val, probe + 1, prev, u);
}

// lower_bound over [lo,hi) started at an arbitrary position h
template <class V>
RBL_HD int64_t pav_lower_bound_near(const V& val, int64_t lo, int64_t hi, double u, int64_t h) {
    if (hi <= lo) return lo;
    if (h < lo) h = lo;
    if (h > hi - 1) h = hi - 1;
    if (val(h) >= u) return pav_gallop_lower_from_hi(val, lo, h + 1, u);  // everything from h on is >= u as well
    int64_t step = 1, prev = h + 1, probe = h + 1;                        // invariant: everything before prev is < u
    while (probe < hi && val(probe) < u) {
        prev = probe + 1;
        probe = h + 1 + step;
        step <<= 1;
    }
    if (probe > hi) probe = hi;
    return pav_lower_bound(val, prev, probe, u);
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

// end of the run of values <= u that contains position p (val(p) <= u): first i in (p, limit) with
// val(i) > u, found by galloping so that the usual short run costs one or two reads
template <class V>
RBL_HD int64_t pav_run_end(const V& val, int64_t p, int64_t limit, double u) {
    int64_t prev = p, step = 1, probe = p + 1;
    while (probe < limit && val(probe) <= u) {
        prev = probe;
        step <<= 1;
        probe = prev + step;
    }
    if (probe > limit) probe = limit;
    return pav_upper_bound(val, prev + 1, probe, u);
}

// start of the run of values >= u that contains position p (val(p) >= u): first i in [limit, p] with
// val(i) >= u, galloping backwards
template <class V>
RBL_HD int64_t pav_run_start(const V& val, int64_t p, int64_t limit, double u) {
    int64_t prev = p, step = 1, probe = p - 1;
    while (probe >= limit && val(probe) >= u) {
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
template <class V, class PS, class PM>
RBL_HD bool pav_merge_search(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a,
                             int64_t b, int64_t c, int64_t* lo_out, int64_t* hi_out, double* v_out) {
    if (!(val(b - 1) > val(b))) return false;
    // ---- left side: first p in [a,b) with Phi(val(p)+) >= 0; val(b-1) always qualifies.
    // Probes gallop outward from the boundary (most pooled blocks are short), then bisect.  A probe
    // decides for the whole run of equal values around it (an earlier pooled block, possibly huge),
    // so the interval jumps to the run's start / end instead of stepping through it.
    int64_t lo = a, hi = b - 1;
    int64_t xlo = b, xhi = c;  // window of upper_bound(val(b..c), val(p)) for the remaining probes
    bool gallop = true;
    int64_t step = 1;
    while (lo < hi) {
        int64_t mid;
        if (gallop) {
            mid = hi - step;
            if (mid < lo) mid = lo;
        } else {
            mid = lo + ((hi - lo) >> 1);
        }
        const double u = val(mid);
        const int64_t l = pav_run_end(val, mid, b, u);  // {L: val > u} = [l, b)
        const int64_t r = gallop ? pav_gallop_upper_from_lo(val, xlo, xhi, u)
                                 : pav_upper_bound(val, xlo, xhi, u);  // {R: val <= u} = [b, r)
        if (pav_phi(loss, rho, ps, pm, l, r, u, +1) >= 0.0) {
            hi = pav_run_start(val, mid, lo, u);
            xhi = r;
            step <<= 1;
        } else {
            lo = l;
            xlo = r;
            gallop = false;
        }
    }
    int64_t lo_star = lo;
    // ---- right side: first p in (b,c) with Phi(val(p)-) > 0; val(b) never qualifies
    lo = b + 1;
    hi = c;
    xlo = a;
    xhi = b;  // window of lower_bound(val(a..b), val(p))
    gallop = true;
    step = 1;
    while (lo < hi) {
        int64_t mid;
        if (gallop) {
            mid = lo + step - 1;
            if (mid > hi - 1) mid = hi - 1;
        } else {
            mid = lo + ((hi - lo) >> 1);
        }
        const double u = val(mid);
        const int64_t l = gallop ? pav_gallop_lower_from_hi(val, xlo, xhi, u)
                                 : pav_lower_bound(val, xlo, xhi, u);  // {L: val >= u} = [l, b)
        const int64_t r = pav_run_start(val, mid, b, u);               // {R: val < u} = [b, r)
        if (pav_phi(loss, rho, ps, pm, l, r, u, -1) > 0.0) {
            hi = r;
            xhi = l;
            gallop = false;
        } else {
            lo = pav_run_end(val, mid, hi, u);
            xlo = l;
            step <<= 1;
        }
    }
    int64_t hi_star = lo;
    // snap to whole runs of equal values (probes inside a run give the same answer, but rounding
    // in Phi near a tie may split one; equal values pool for free)
    lo_star = pav_run_start(val, lo_star, a, val(lo_star));
    hi_star = pav_run_end(val, hi_star - 1, c, val(hi_star - 1));
    double v = pav_block_value(loss, rho, ps, pm, lo_star, hi_star);
    // exact arithmetic guarantees val(lo*-1) <= v <= val(hi*); clamp so rounding can never break the
    // "non-decreasing inside a solved range" invariant the bound searches rely on
    if (lo_star > a && v < val(lo_star - 1)) v = val(lo_star - 1);
    if (hi_star < c && v > val(hi_star)) v = val(hi_star);
    *lo_out = lo_star;
    *hi_out = hi_star;
    *v_out = v;
    return true;
}

// =================================================================================================
// Global levels (ranges wider than one chunk, values in global memory): the same merge as
// pav_merge_search, but with 32-ary searches — the 32 lanes of a warp probe 32 positions per round,
// each lane running its own opposite-side bound search, so the dependent-load chain of a merge is
// ~3 rounds instead of ~20 bisection steps.  Round 0 probes exponentially growing distances from the
// boundary (most pooled blocks are short), later rounds subdivide the bracket evenly.
// The per-lane probes and the round bookkeeping below are shared between the CUDA kernel (one lane
// per probe, ballot/shuffle) and the host emulation in tests/native (loops over lanes).
// =================================================================================================

// left-side probe at p in [a,b): the whole run of equal values around p joins the pooled block iff
// Phi(val(p)+) >= 0.  [xlo,xhi] bounds upper_bound(val[b..c), val(p)); outputs: that bound, and the run.
template <class V, class PS, class PM>
RBL_HD bool pav_probe_left(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a, int64_t b,
                           int64_t p, int64_t xlo, int64_t xhi, bool gallop, int64_t* r_out, int64_t* rs_out,
                           int64_t* re_out) {
    const double u = val(p);
    const int64_t l = pav_run_end(val, p, b, u);                                // {L: val > u} = [l, b)
    const int64_t r = gallop ? pav_gallop_upper_from_lo(val, xlo, xhi, u)
                             : pav_upper_bound(val, xlo, xhi, u);                // {R: val <= u} = [b, r)
    *r_out = r;
    *re_out = l;
    *rs_out = pav_run_start(val, p, a, u);
    return pav_phi(loss, rho, ps, pm, l, r, u, +1) >= 0.0;
}

// right-side probe at p in (b,c): the run around p stays out of the pooled block iff Phi(val(p)-) > 0
template <class V, class PS, class PM>
RBL_HD bool pav_probe_right(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t b, int64_t c,
                            int64_t p, int64_t xlo, int64_t xhi, bool gallop, int64_t* l_out, int64_t* rs_out,
                            int64_t* re_out) {
    const double u = val(p);
    const int64_t l = gallop ? pav_gallop_lower_from_hi(val, xlo, xhi, u)
                             : pav_lower_bound(val, xlo, xhi, u);                // {L: val >= u} = [l, b)
    const int64_t r = pav_run_start(val, p, b, u);                               // {R: val < u} = [b, r)
    *l_out = l;
    *rs_out = r;
    *re_out = pav_run_end(val, p, c, u);
    return pav_phi(loss, rho, ps, pm, l, r, u, -1) > 0.0;
}

// the same probes with the opposite-side bound searched outward from a guess h2 of it (warm start: the pooled
// block of the previous z-step — ranks move little between ADMM iterations)
template <class V, class PS, class PM>
RBL_HD bool pav_probe_left_near(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a, int64_t b,
                                int64_t c, int64_t p, int64_t h2, int64_t* r_out, int64_t* rs_out, int64_t* re_out) {
    const double u = val(p);
    const int64_t l = pav_run_end(val, p, b, u);
    const int64_t r = pav_upper_bound_near(val, b, c, u, h2);
    *r_out = r;
    *re_out = l;
    *rs_out = pav_run_start(val, p, a, u);
    return pav_phi(loss, rho, ps, pm, l, r, u, +1) >= 0.0;
}

template <class V, class PS, class PM>
RBL_HD bool pav_probe_right_near(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a, int64_t b,
                                 int64_t c, int64_t p, int64_t h2, int64_t* l_out, int64_t* rs_out, int64_t* re_out) {
    const double u = val(p);
    const int64_t l = pav_lower_bound_near(val, a, b, u, h2);
    const int64_t r = pav_run_start(val, p, b, u);
    *l_out = l;
    *rs_out = r;
    *re_out = pav_run_end(val, p, c, u);
    return pav_phi(loss, rho, ps, pm, l, r, u, -1) > 0.0;
}

// probe position of lane j of a WARM-STARTED first round: 32 positions spaced 32 ranks apart around the guess h
// (h - 480 .. h + 512), clamped to [lo, hi] (ascending in j; duplicates are harmless).  A block end that moved by
// fewer than ~500 ranks since the guess was taken is bracketed to 32 candidates, which the next round resolves
// exactly; all these probes stay inside the shared-memory windows the device kernel keeps around the guesses.
// Farther moves fall back to the even 32-ary subdivision of what is left.
#define RBL_HINT_STRIDE 32
// stride: RBL_HINT_STRIDE, or a finer spacing when the last guesses were close (RBL_HINT_STRIDE_NEAR: the 32 probes
// then cover h - 120 .. h + 128, whose lookups on the OPPOSITE side also stay inside that side's window).
#define RBL_HINT_STRIDE_NEAR 8
RBL_HD int64_t pav_hint_pos(int64_t h, int64_t lo, int64_t hi, int j, int stride = RBL_HINT_STRIDE) {
    int64_t p = h + (int64_t)(j - 15) * stride;
    if (p < lo) p = lo;
    if (p > hi) p = hi;
    return p;
}

// probe position of lane j when the unknown candidates are [lo, lo+width); lanes j >= width idle if width <= 32
RBL_HD int64_t pav_kary_pos(int64_t lo, int64_t width, int j) {
    if (width <= 32) return lo + j;
    return lo + (int64_t)(((long long)(j + 1) * (long long)width) / 33);
}

template <class V, class PS, class PM>
RBL_HD void pav_kary_finish(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a, int64_t c,
                            int64_t lo_star, int64_t hi_star, int64_t* lo_out, int64_t* hi_out, double* v_out) {
    lo_star = pav_run_start(val, lo_star, a, val(lo_star));
    hi_star = pav_run_end(val, hi_star - 1, c, val(hi_star - 1));
    double v = pav_block_value(loss, rho, ps, pm, lo_star, hi_star);
    if (lo_star > a) {
        const double vl = val(lo_star - 1);
        if (v < vl) v = vl;
    }
    if (hi_star < c) {
        const double vr = val(hi_star);
        if (v > vr) v = vr;
    }
    *lo_out = lo_star;
    *hi_out = hi_star;
    *v_out = v;
}

// lane-loop version (host emulation); the CUDA kernel mirrors it with one lane per probe.
// hint_lo / hint_hi (-1: none): the pooled block [lo*, hi*) this merge produced at the previous z-step.  With hints
// the first round of each side probes around the guess (pav_hint_pos) instead of at exponential distances from the
// boundary, and its opposite-side bounds are searched outward from the other guess; the result never depends on them.
template <class V, class PS, class PM>
RBL_HD bool pav_merge_search_kary(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a,
                                  int64_t b, int64_t c, int64_t* lo_out, int64_t* hi_out, double* v_out,
                                  int64_t hint_lo = -1, int64_t hint_hi = -1, int hint_stride = RBL_HINT_STRIDE) {
    if (!(val(b - 1) > val(b))) return false;
    const bool hinted = hint_lo >= a && hint_lo < b && hint_hi > b && hint_hi <= c;
    // ---- left: first p in [a, b-1] whose probe is true (b-1 is): pattern over p is F..F T..T
    int64_t lo = a, hi = b - 1, xlo = b, xhi = c;
    bool first = true;
    if (hinted && lo < hi) {
        int64_t rr[32], rs[32], re[32];
        bool pr[32];
        for (int j = 0; j < 32; ++j)
            pr[j] = pav_probe_left_near(loss, rho, val, ps, pm, a, b, c, pav_hint_pos(hint_lo, lo, hi, j, hint_stride), hint_hi,
                                        &rr[j], &rs[j], &re[j]);
        int f = -1;
        for (int j = 0; j < 32; ++j)
            if (pr[j]) { f = j; break; }
        if (f >= 0) {
            hi = rs[f] < hi ? rs[f] : hi;
            xhi = rr[f];
            if (f > 0) { lo = re[f - 1] > lo ? re[f - 1] : lo; xlo = rr[f - 1]; }
        } else {
            lo = re[31] > lo ? re[31] : lo;
            xlo = rr[31];
        }
        if (hi < lo) hi = lo;
        first = false;
    }
    while (lo < hi) {
        const int64_t width = hi - lo;
        int active;
        int64_t pos[32], rr[32], rs[32], re[32];
        bool pr[32];
        if (first) {  // exponential distances hi - 2^j: lanes see T..T F..F
            active = 0;
            while (active < 32 && ((int64_t)1 << active) <= width) ++active;
            for (int j = 0; j < active; ++j) pos[j] = hi - ((int64_t)1 << j);
        } else {
            active = width <= 32 ? (int)width : 32;
            for (int j = 0; j < active; ++j) pos[j] = pav_kary_pos(lo, width, j);
        }
        for (int j = 0; j < active; ++j)
            pr[j] = pav_probe_left(loss, rho, val, ps, pm, a, b, pos[j], xlo, xhi, first, &rr[j], &rs[j], &re[j]);
        if (first) {
            int f = active;  // first false lane
            for (int j = 0; j < active; ++j)
                if (!pr[j]) { f = j; break; }
            if (f < active) { lo = re[f] > lo ? re[f] : lo; xlo = rr[f]; }
            if (f > 0) { hi = rs[f - 1] < hi ? rs[f - 1] : hi; xhi = rr[f - 1]; }
            first = false;
        } else {
            int f = -1;  // first true lane
            for (int j = 0; j < active; ++j)
                if (pr[j]) { f = j; break; }
            if (f >= 0) {
                hi = rs[f] < hi ? rs[f] : hi;
                xhi = rr[f];
                if (f > 0) { lo = re[f - 1] > lo ? re[f - 1] : lo; xlo = rr[f - 1]; }
            } else {
                lo = re[active - 1] > lo ? re[active - 1] : lo;
                xlo = rr[active - 1];
            }
        }
        if (hi < lo) hi = lo;
    }
    const int64_t lo_star = lo;
    // ---- right: first p in [b+1, c) whose probe is true, else c: pattern over p is F..F T..T
    lo = b + 1;
    hi = c;
    xlo = a;
    xhi = b;
    first = true;
    if (hinted && lo < hi) {
        int64_t ll[32], rs[32], re[32];
        bool pr[32];
        for (int j = 0; j < 32; ++j)
            pr[j] = pav_probe_right_near(loss, rho, val, ps, pm, a, b, c, pav_hint_pos(hint_hi, lo, hi - 1, j, hint_stride), hint_lo,
                                         &ll[j], &rs[j], &re[j]);
        int f = -1;
        for (int j = 0; j < 32; ++j)
            if (pr[j]) { f = j; break; }
        if (f >= 0) {
            hi = rs[f] < hi ? rs[f] : hi;
            xhi = ll[f];
            if (f > 0) { lo = re[f - 1] > lo ? re[f - 1] : lo; xlo = ll[f - 1]; }
        } else {
            lo = re[31] > lo ? re[31] : lo;
            xlo = ll[31];
        }
        if (hi < lo) hi = lo;
        first = false;
    }
    while (lo < hi) {
        const int64_t width = hi - lo;
        int active;
        int64_t pos[32], ll[32], rs[32], re[32];
        bool pr[32];
        if (first) {  // exponential distances lo + 2^j - 1
            active = 0;
            while (active < 32 && ((int64_t)1 << active) <= width) ++active;
            for (int j = 0; j < active; ++j) pos[j] = lo + ((int64_t)1 << j) - 1;
        } else {
            active = width <= 32 ? (int)width : 32;
            for (int j = 0; j < active; ++j) pos[j] = pav_kary_pos(lo, width, j);
        }
        for (int j = 0; j < active; ++j)
            pr[j] = pav_probe_right(loss, rho, val, ps, pm, b, c, pos[j], xlo, xhi, first, &ll[j], &rs[j], &re[j]);
        int f = -1;  // first true lane
        for (int j = 0; j < active; ++j)
            if (pr[j]) { f = j; break; }
        if (f >= 0) {
            hi = rs[f] < hi ? rs[f] : hi;
            xhi = ll[f];
            if (f > 0) { lo = re[f - 1] > lo ? re[f - 1] : lo; xlo = ll[f - 1]; }
        } else {
            lo = re[active - 1] > lo ? re[active - 1] : lo;
            xlo = ll[active - 1];
        }
        first = false;
        if (hi < lo) hi = lo;
    }
    pav_kary_finish(loss, rho, val, ps, pm, a, c, lo_star, lo, lo_out, hi_out, v_out);
    return true;
}
