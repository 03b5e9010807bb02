// prox_core.h — scalar building blocks of the rank-based z-step, shared by the CUDA kernels and
// by the host-side emulation harness in tests/native/ (compiled with gcc, no GPU needed).
//
// Replaces (semantics): src/util/individual_solver.py:44-130 (per-element / per-block prox),
// src/util/pav.py:134-146 (block value from the block means), src/util/PAV_cpt.py:41-94.
//
// Everything here is branch-light fp64 and uses only +,-,*,/,exp so host and device agree to
// rounding.  No fused contraction is relied upon for correctness.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RBL_HD __host__ __device__ __forceinline__
#define RBL_HDM __host__ __device__ __forceinline__
#else
#define RBL_HD static inline
#define RBL_HDM inline
#endif

#define RBL_LOSS_BCE 0
#define RBL_LOSS_HINGE 1

// ---- double-double arithmetic (error-free transformations) ------------------------------------
// Prefix sums of sigma and of the sorted margins are kept as unevaluated sums hi+lo so that block
// sums obtained by subtraction, S[l,r) = P[r] - P[l], are accurate to ~1e-30 relative to |P| —
// i.e. as good as summing the block directly (the reference uses np.mean / running sums per block,
// pav.py:26-27,134-140).
struct dd_t {
    double hi, lo;
};

#if defined(__CUDA_ARCH__)
#define RBL_ADD(a, b) __dadd_rn((a), (b))
#define RBL_SUB(a, b) __dsub_rn((a), (b))
#define RBL_MUL(a, b) __dmul_rn((a), (b))
#else
#define RBL_ADD(a, b) ((a) + (b))
#define RBL_SUB(a, b) ((a) - (b))
#define RBL_MUL(a, b) ((a) * (b))
#endif

RBL_HD dd_t dd_make(double x) {
    dd_t r;
    r.hi = x;
    r.lo = 0.0;
    return r;
}

RBL_HD dd_t dd_two_sum(double a, double b) {
    dd_t r;
    double s = RBL_ADD(a, b);
    double bb = RBL_SUB(s, a);
    double e = RBL_ADD(RBL_SUB(a, RBL_SUB(s, bb)), RBL_SUB(b, bb));
    r.hi = s;
    r.lo = e;
    return r;
}

RBL_HD dd_t dd_add(dd_t a, dd_t b) {
    dd_t s = dd_two_sum(a.hi, b.hi);
    double lo = RBL_ADD(s.lo, RBL_ADD(a.lo, b.lo));
    // renormalise (fast two-sum, |s.hi| >= |lo|)
    double hi = RBL_ADD(s.hi, lo);
    dd_t r;
    r.lo = RBL_SUB(lo, RBL_SUB(hi, s.hi));
    r.hi = hi;
    return r;
}

RBL_HD dd_t dd_add_d(dd_t a, double b) { return dd_add(a, dd_make(b)); }

RBL_HD dd_t dd_neg(dd_t a) {
    dd_t r;
    r.hi = -a.hi;
    r.lo = -a.lo;
    return r;
}

// (a - b) rounded to double
RBL_HD double dd_diff(dd_t a, dd_t b) {
    dd_t r = dd_add(a, dd_neg(b));
    return r.hi + r.lo;
}

// ---- losses -------------------------------------------------------------------------------------
// sigmoid(x) = e^x / (1 + e^x), stable on both sides (individual_solver.py:44-49 safe_1divexp)
RBL_HD double rbl_sigmoid(double x) {
    double e = exp(-fabs(x));
    double s = 1.0 / (1.0 + e);
    return x >= 0.0 ? s : e * s;
}

// log(1 + e^x), stable (individual_solver.py:52-57 log1exp)
RBL_HD double rbl_log1pexp(double x) {
    double l = log1p(exp(-fabs(x)));
    return x > 0.0 ? x + l : l;
}

RBL_HD double rbl_margin_loss(int loss, double u) {
    if (loss == RBL_LOSS_HINGE) return fmax(1.0 + u, 0.0);  // objective.py:23-24 in terms of u = -y x.w
    return rbl_log1pexp(u);                                  // objective.py:11-16
}

// ---- block prox ---------------------------------------------------------------------------------
// minimiser over z of  sbar*loss(z) + rho/2 (z - mbar)^2  with sbar, mbar the block MEANS
// (pav.py:134-140 passes sigma_param/len, m_param/len to individual_solver).
//
// BCE: root of g(z) = sbar*sigmoid(z) + rho (z - mbar), g increasing, root in [mbar - sbar/rho, mbar].
//   sigmoid is convex on z<0 and concave on z>0; Newton started on the side where g*g'' >= 0
//   converges monotonically, so no line search is needed (the reference needs its Armijo damping,
//   individual_solver.py:96-101, because undamped Newton from z = m cycles when sbar/rho >> 1):
//     root <  0 (g(0) > 0): start at min(mbar, 0);  root >= 0: start at max(mbar - sbar/rho, 0).
//   A bracket with bisection fallback covers the rounding-level end game; solved to machine eps
//   (the reference stops at ||delta||_2 < 1e-6 over the whole vector, :103).
// hinge: closed form (the reference bisects 50 times with a global early exit, :15-42 — waived).
RBL_HD double rbl_block_prox(int loss, double sbar, double mbar, double rho) {
    if (loss == RBL_LOSS_HINGE) {
        if (mbar < -1.0) return mbar;
        double c = mbar - sbar / rho;
        return c > -1.0 ? c : -1.0;
    }
    if (!(sbar > 0.0)) return mbar;
    double lo = mbar - sbar / rho, hi = mbar;
    double z;
    if (0.5 * sbar - rho * mbar > 0.0) {
        z = mbar < 0.0 ? mbar : 0.0;
        if (hi > 0.0) hi = 0.0;
    } else {
        z = lo > 0.0 ? lo : 0.0;
        if (lo < 0.0) lo = 0.0;
    }
    // g is evaluated as sbar*sigmoid(z) + rho*(z - mbar): its rounding error is ~eps*max(|z|,|mbar|)*rho,
    // so the root is only defined to ~eps*max(|z|,|mbar|).  Stopping on that (not on eps*|z|) matters:
    // roots near zero would otherwise jitter for the full iteration budget and stall their warp.
    const double floor_abs = 2.3e-16 * fabs(mbar);
    double zprev = lo - 1.0;  // outside the bracket: never equal to an iterate
    for (int it = 0; it < 64; ++it) {
        double s = rbl_sigmoid(z);
        double g = sbar * s + rho * (z - mbar);
        if (g == 0.0) break;
        if (g > 0.0) hi = z; else lo = z;
        double dg = sbar * s * (1.0 - s) + rho;
        double zn = z - g / dg;
        if (!(zn >= lo && zn <= hi)) zn = 0.5 * (lo + hi);
        // zn == zprev: the iteration has collapsed onto two neighbouring doubles with residuals of opposite
        // sign and would bounce between them for the whole budget (0.5% of the elements did, and each one
        // stalled its warp for 64 rounds)
        if (zn == z || zn == zprev) break;
        double dz = fabs(zn - z);
        zprev = z;
        z = zn;
        if (dz <= fmax(2.3e-16 * fabs(z), floor_abs)) break;
    }
    return z;
}

// one-sided derivatives of the loss at u (side < 0: left, side > 0: right); equal for BCE
RBL_HD double rbl_loss_deriv(int loss, double u, int side) {
    if (loss == RBL_LOSS_HINGE) return (side < 0 ? (u > -1.0) : (u >= -1.0)) ? 1.0 : 0.0;
    return rbl_sigmoid(u);
}

// ---- radix-sort key transform (algorithms.py:92-93 sorts fp64 ascending; NaN last, -0.0 == +0.0)
RBL_HD uint64_t rbl_key_from_bits(uint64_t bits) {
    if ((bits << 1) == 0ull) bits = 0ull;                                   // -0.0 -> +0.0
    if ((bits & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) bits = 0x7ff8000000000000ull;  // NaN last
    return bits ^ ((bits >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}

RBL_HD uint64_t rbl_bits_from_key(uint64_t key) {
    return key ^ ((key >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull);
}
