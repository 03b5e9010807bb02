/*
 * oracle/pav_oracle.c — CPU restatement of the reference's rank-based z-step prox.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (admm-for-rank-based-loss_b200/) never does.
 *
 * What it restates (semantics, not structure):
 *   - src/util/individual_solver.py:90-130  per-element / per-block prox
 *       BCE   : root of  sbar*sigmoid(z) + rho*(z - mbar) = 0   (reference: damped vector
 *               Newton, global stop |delta|_2 < 1e-6; here: bracketed Newton to machine eps)
 *       hinge : closed form of the same stationarity condition (reference: 50-step vector
 *               bisection with a global early exit, individual_solver.py:15-42 — waived,
 *               see SURVEY.md §8a(iii))
 *   - src/util/pav.py:93-178  pool-adjacent-violators over the sorted margins.  The
 *     reference sweeps (merge every violating run, re-solve, repeat); the fixed point is
 *     the unique isotonic prox, which a classic stack PAV reaches in O(n).  Block value
 *     depends only on (sum sigma, sum m, count)  (pav.py:134-140: means of the block).
 *   - src/util/PAV_cpt.py:203-293  EHRM: min(B, isotonic prox with sigma = a) or max(B, isotonic
 *     prox with sigma = b), chosen by the scalar comparison of :222-226 (made in
 *     rbl_oracle.py::ehrm_pav); rbl_oracle_pav() takes the optional lower clip.
 *   - scikit-learn's Lasso coordinate descent (third-party, un-vendored: README.md pins
 *     scikit-learn 1.2.2), called by the small-problem w-step at src/optim/algorithms.py:194-197:
 *     rbl_oracle_lasso_cd() restates sklearn/linear_model/_cd_fast.pyx::enet_coordinate_descent
 *     of that release (cyclic, no screening, duality-gap stop).
 *
 * Parity status: the reference holds no golden vectors for this path ("parity unpinned"
 * by the reference).  This file is pinned instead against outputs of the shimmed
 * reference itself, committed under tests/golden/ (oracle/gen_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define RBL_LOSS_BCE 0
#define RBL_LOSS_HINGE 1

static double sigmoid_stable(double x) { /* individual_solver.py:44-49 safe_1divexp */
    if (x > 0) return 1.0 / (1.0 + exp(-x));
    double e = exp(x);
    return e / (1.0 + e);
}

/* minimiser of  sbar*loss(z) + rho/2 (z - mbar)^2 */
double rbl_oracle_prox(int loss, double sbar, double mbar, double rho) {
    if (loss == RBL_LOSS_HINGE) {
        /* loss(u) = max(0, 1+u): derivative 0 for u<-1, 1 for u>-1 */
        if (mbar < -1.0) return mbar;
        double c = mbar - sbar / rho;
        if (c > -1.0) return c;
        return -1.0;
    }
    if (sbar == 0.0) return mbar;
    /* g(z) = sbar*sigmoid(z) + rho (z - mbar) is increasing with g(mbar - sbar/rho) <= 0 <= g(mbar).
     * sigmoid is convex on z<0 and concave on z>0, so Newton started where g*g'' >= 0 converges
     * monotonically (undamped Newton from z = mbar can 2-cycle when sbar/rho >> 1):
     *   root < 0  (g(0) > 0): start at min(mbar, 0)            (g >= 0, convex side)
     *   root >= 0 (g(0) <= 0): start at max(mbar - sbar/rho, 0) (g <= 0, concave side)
     * A bracket [lo, hi] with bisection fallback guards the rounding-level end game. */
    double lo = mbar - sbar / rho, hi = mbar;
    double g0 = 0.5 * sbar - rho * mbar;
    double z;
    if (g0 > 0) { z = mbar < 0 ? mbar : 0.0; if (hi > 0) hi = 0.0; }
    else        { z = lo > 0 ? lo : 0.0;     if (lo < 0) lo = 0.0; }
    for (int it = 0; it < 100; ++it) {
        double s = sigmoid_stable(z);
        double g = sbar * s + rho * (z - mbar);
        if (g == 0.0) break;
        if (g > 0) hi = z; else lo = z;
        double dg = sbar * s * (1.0 - s) + rho;
        double zn = z - g / dg;
        if (!(zn >= lo && zn <= hi)) zn = 0.5 * (lo + hi);
        if (zn == z) break;
        double dz = fabs(zn - z);
        z = zn;
        if (dz <= 2.3e-16 * fabs(z)) break;
    }
    return z;
}

/*
 * Exact isotonic prox on sorted margins:
 *   argmin_{z1<=...<=zn} sum_i sigma_i*loss(z_i) + rho/2 (z_i - m_i)^2
 * followed by z = max(z, clip) if use_clip.  Returns the number of blocks.
 */
int64_t rbl_oracle_pav(int loss, int64_t n, const double* sigma, const double* m, double rho,
                       int use_clip, double clip, double* z_out) {
    if (n <= 0) return 0;
    long double* ss = (long double*)malloc(sizeof(long double) * (size_t)n);
    long double* sm = (long double*)malloc(sizeof(long double) * (size_t)n);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    double* val = (double*)malloc(sizeof(double) * (size_t)n);
    int64_t top = 0;
    for (int64_t i = 0; i < n; ++i) {
        ss[top] = sigma[i]; sm[top] = m[i]; cnt[top] = 1;
        val[top] = rbl_oracle_prox(loss, sigma[i], m[i], rho);
        ++top;
        while (top > 1 && val[top - 2] > val[top - 1]) {
            ss[top - 2] += ss[top - 1]; sm[top - 2] += sm[top - 1]; cnt[top - 2] += cnt[top - 1];
            --top;
            val[top - 1] = rbl_oracle_prox(loss, (double)(ss[top - 1] / cnt[top - 1]),
                                           (double)(sm[top - 1] / cnt[top - 1]), rho);
        }
    }
    int64_t pos = 0;
    for (int64_t b = 0; b < top; ++b) {
        double v = val[b];
        if (use_clip && v < clip) v = clip;
        for (int64_t k = 0; k < cnt[b]; ++k) z_out[pos++] = v;
    }
    free(ss); free(sm); free(cnt); free(val);
    return top;
}

/* element-wise prox (no pooling): the reference's PAV_solver.__init__ stage, pav.py:63 */
void rbl_oracle_prox_vec(int loss, int64_t n, const double* sigma, const double* m, double rho,
                         double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = rbl_oracle_prox(loss, sigma[i], m[i], rho);
}

/*
 * scikit-learn 1.2.2 `enet_coordinate_descent` (sklearn/linear_model/_cd_fast.pyx) for
 *     min_w  1/2 ||y - X w||^2 + alpha ||w||_1            (beta = 0: Lasso, l1_ratio = 1)
 * as reached from Lasso(alpha=a, tol=1e-8, fit_intercept=False, max_iter=50000).fit(X, y)
 * (algorithms.py:195-196): alpha = a * n_samples, tol is scaled by y.y, w starts at 0 (a fresh
 * estimator has no coef_ to warm-start from), cyclic coordinate order, residual R kept up to date by
 * axpy, and after every sweep whose largest update is below tol relative to the largest |w_j| the
 * duality gap decides.  X is column-major n x d.  Returns the number of sweeps.
 */
int rbl_oracle_lasso_cd(int n, int d, const double* X, const double* y, double alpha, double tol,
                        int max_iter, double* w, double* gap_out) {
    double* R = (double*)malloc(sizeof(double) * (size_t)n);
    double* nrm = (double*)malloc(sizeof(double) * (size_t)d);
    double yy = 0.0, gap = tol + 1.0;
    const double d_w_tol = tol;
    for (int i = 0; i < n; ++i) { R[i] = y[i]; yy += y[i] * y[i]; }
    for (int j = 0; j < d; ++j) {
        double a = 0.0;
        const double* xj = X + (size_t)j * n;
        for (int i = 0; i < n; ++i) a += xj[i] * xj[i];
        nrm[j] = a;
        w[j] = 0.0;
    }
    tol *= yy;
    int it = 0;
    for (it = 0; it < max_iter; ++it) {
        double w_max = 0.0, d_w_max = 0.0;
        for (int j = 0; j < d; ++j) {
            if (nrm[j] == 0.0) continue;
            const double* xj = X + (size_t)j * n;
            const double w_j = w[j];
            if (w_j != 0.0)
                for (int i = 0; i < n; ++i) R[i] += w_j * xj[i];
            double tmp = 0.0;
            for (int i = 0; i < n; ++i) tmp += xj[i] * R[i];
            const double a = fabs(tmp) - alpha;
            w[j] = (tmp > 0 ? 1.0 : (tmp < 0 ? -1.0 : 0.0)) * (a > 0 ? a : 0.0) / nrm[j];
            if (w[j] != 0.0)
                for (int i = 0; i < n; ++i) R[i] -= w[j] * xj[i];
            const double d_w_j = fabs(w[j] - w_j);
            if (d_w_j > d_w_max) d_w_max = d_w_j;
            if (fabs(w[j]) > w_max) w_max = fabs(w[j]);
        }
        if (w_max == 0.0 || d_w_max / w_max < d_w_tol || it == max_iter - 1) {
            double dual_norm = 0.0, R2 = 0.0, Ry = 0.0, l1 = 0.0;
            for (int j = 0; j < d; ++j) {
                const double* xj = X + (size_t)j * n;
                double a = 0.0;
                for (int i = 0; i < n; ++i) a += xj[i] * R[i];
                if (fabs(a) > dual_norm) dual_norm = fabs(a);
                l1 += fabs(w[j]);
            }
            for (int i = 0; i < n; ++i) { R2 += R[i] * R[i]; Ry += R[i] * y[i]; }
            double c;
            if (dual_norm > alpha) {
                c = alpha / dual_norm;
                gap = 0.5 * (R2 + R2 * c * c);
            } else {
                c = 1.0;
                gap = R2;
            }
            gap += alpha * l1 - c * Ry;
            if (gap < tol) { ++it; break; }
        }
    }
    if (gap_out) *gap_out = gap;
    free(R); free(nrm);
    return it;
}
