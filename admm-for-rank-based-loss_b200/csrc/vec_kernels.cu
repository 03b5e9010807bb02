// vec_kernels.cu — n- and d-vector kernels of the ADMM loop: everything that is not a pass over D,
// a sort or the PAV.  All reductions are two-stage with a fixed grid, so results are bit-reproducible
// run to run and identical on every rank of a row-sharded job.
//
// Replaces: src/optim/algorithms.py:23 (D = -y*X), :89 (margins), :103-104 (scatter), :132-136 (dual
// update + residual norms), src/util/fast_lasso.py:44-65 (FISTA line search / momentum / stop test —
// here a device-resident state machine), src/optim/objective.py:71-87 (rank-weighted objective).
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kVecThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// sum over the block, result valid in every thread; sh must hold >= 33 doubles
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// ---- D = -y (.) X with zero-filled padding columns (algorithms.py:23) ---------------------------
template <typename T>
__global__ void build_design_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ y,
                                    T* __restrict__ D, int64_t ld, int64_t n, int d) {
    const int64_t total = n * ld;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / ld;
        const int c = (int)(idx - i * ld);
        D[idx] = (T)((c < d) ? -(y[i] * X[i * ldx + c]) : 0.0);  // fp32 storage: one rounding of the fp64 product
    }
}

// ---- per-CTA partials of a pass -> red[0..d) = D^T r, red[d] = ||r||^2, red[d+1] = c0 -------------
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ gpart,
                                                              const double* __restrict__ sspart,
                                                              const double* __restrict__ c0part, int nparts, int nc0,
                                                              int64_t ld, int d, double* __restrict__ red,
                                                              const FistaState* st) {
    rbl_pdl_wait();
    // 32 columns x 8 row-slices per CTA: coalesced 256-byte row segments, 8 independent partial sums per
    // column combined in a fixed order (slice 0..7) => deterministic
    __shared__ double sh[8][33];
    if (st && st->done) return;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    double a = 0.0;
    if (c < d)
        for (int k = ry; k < nparts; k += 8) a += gpart[(size_t)k * ld + c];
    sh[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && c < d) {
        double t = sh[0][cx];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += sh[q][cx];
        red[c] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double s = 0.0;
        for (int k = lane; k < nparts; k += 32) s += sspart[k];
        s = warp_sum(s);
        double c0 = 0.0;
        if (c0part) {
            for (int k = lane; k < nc0; k += 32) c0 += c0part[k];
            c0 = warp_sum(c0);
        }
        if (lane == 0) {
            red[d] = s;
            red[d + 1] = c0;
        }
    }
}

// ---- FISTA control flow on the device (fast_lasso.py:40-67) -------------------------------------
// One CTA.  Consumes red = [D^T r(beta), ||r(beta)||^2, c0] of the pass that just ran at `beta`
// and either (init) seeds the iteration, (reject) grows L and forms the next trial, or (accept)
// applies momentum, tests convergence and forms the first trial of the next iteration.
// The gradient and ||.||^2 at the extrapolated point beta_p are NOT recomputed with passes over D:
// both are affine in beta, so  g(beta_p) = g(beta) + t1 (g(beta) - g(beta_prev))  and
// r(beta_p) = r(beta) + t1 (r(beta) - r(beta_prev))  (combine kernel) — one D pass per trial instead of
// the reference's three matvecs per iteration, same iterates up to rounding.
__global__ void __launch_bounds__(1024) fista_update_kernel(FistaState* st, int d, const double* __restrict__ red,
                                                            int64_t red_stride, double* beta, double* beta_p,
                                                            double* beta_prev, double* g_p, double* g_prev,
                                                            int64_t vstride, const float* __restrict__ pow_tab) {
    __shared__ double sh[33];
    {   // batched mode: one CTA per instance
        const int64_t g = blockIdx.x;
        st += g;
        red += g * red_stride;
        beta += g * vstride;
        beta_p += g * vstride;
        beta_prev += g * vstride;
        g_p += g * vstride;
        g_prev += g * vstride;
    }
    if (st->done) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const FistaState S = *st;
    const double ss = red[d];
    float L_prev = S.L_prev, L_cur = S.L_cur;
    int i_k = S.i_k, k = S.k, cur = S.cur, accepted = 0, c0_from_red = 0, done = 0;
    double t = S.t, c0 = S.c0, t1 = S.t1, crit = S.crit;

    if (k < 0) {
        // the initial pass ran at beta = w (beta_p = beta_prev = w, fast_lasso.py:37-38)
        for (int c = tid; c < d; c += nt) {
            const double bw = beta[c], g = red[c];
            beta_p[c] = bw;
            beta_prev[c] = bw;
            g_p[c] = g;
            g_prev[c] = g;
        }
        c0 = ss;
        k = 0;
        i_k = 0;
        t = 1.0;
        L_cur = __fmul_rn(L_prev, pow_tab[0]);
        accepted = 2;
        cur ^= 1;
    } else {
        if (S.c0_from_red) c0 = red[d + 1];
        const double lhs = ss - c0;
        if (lhs > S.rhs) {  // cond = (LHS > RHS), :56 — false for NaN, like the reference
            ++i_k;
            L_cur = __fmul_rn(L_prev, pow_tab[i_k < 127 ? i_k : 127]);
        } else {
            L_prev = L_cur;
            const double tnext = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;  // :59
            t1 = (t - 1.0) / tnext;                                      // :61
            double a = 0.0;
            for (int c = tid; c < d; c += nt) {
                const double df = beta[c] - beta_prev[c];  // :60
                a = fma(df, df, a);
                beta_p[c] = beta[c] + t1 * df;             // :62
            }
            crit = sqrt(block_sum(a, sh));                 // :63
            ++k;
            if (crit < S.tol || k >= S.max_iter) {
                done = 1;  // result is beta; its residual sits in rbuf[cur]
            } else {
                t = tnext;
                for (int c = tid; c < d; c += nt) {
                    const double gb = red[c];
                    g_p[c] = gb + t1 * (gb - g_prev[c]);
                    g_prev[c] = gb;
                    beta_prev[c] = beta[c];
                }
                i_k = 0;
                L_cur = __fmul_rn(L_prev, pow_tab[0]);
                accepted = 1;
                c0_from_red = 1;
                cur ^= 1;
            }
        }
    }
    double rhs = S.rhs;
    if (!done) {
        __syncthreads();
        // trial: beta = soft(beta_p + g_p / L_cur, lam / L_cur)   (:47-49)
        const double Ld = (double)L_cur;
        const double thr = S.thr_f32 ? (double)__fdiv_rn((float)S.lam, L_cur) : S.lam / Ld;
        double r1 = 0.0, r2 = 0.0;
        for (int c = tid; c < d; c += nt) {
            const double bp = beta_p[c], g = g_p[c];
            const double bs = bp + g / Ld;
            const double mag = fmax(fabs(bs) - thr, 0.0);
            const double sgn = (bs > 0.0) ? 1.0 : ((bs < 0.0) ? -1.0 : 0.0);
            const double bn = mag * sgn;
            beta[c] = bn;
            const double df = bn - bp;
            r1 = fma(df, df, r1);
            r2 = fma(df, g, r2);
        }
        r1 = block_sum(r1, sh);
        r2 = block_sum(r2, sh);
        rhs = Ld * r1 - 2.0 * r2;  // :53
    }
    if (tid == 0) {
        st->t = t;
        st->c0 = c0;
        st->rhs = rhs;
        st->t1 = t1;
        st->crit = crit;
        st->ss_last = ss;
        st->L_prev = L_prev;
        st->L_cur = L_cur;
        st->i_k = i_k;
        st->k = k;
        st->done = done;
        st->passes = S.passes + 1;
        st->trials = S.trials + (S.k < 0 ? 0 : 1);
        st->accepted = accepted;
        st->cur = cur;
        st->c0_from_red = c0_from_red;
    }
}

// after an accepted step: partial sums of || r(beta) + t1 (r(beta) - r(beta_prev)) ||^2 = ||b - D beta_p||^2
__global__ void fista_combine_kernel(const FistaState* st, const double* r0, const double* r1, int64_t n,
                                     double* __restrict__ c0part) {
    __shared__ double sh[33];
    {   // batched mode: blockIdx.y = instance
        const int64_t g = blockIdx.y;
        st += g;
        r0 += g * n;
        r1 += g * n;
        c0part += g * gridDim.x;
    }
    if (st->done || st->accepted != 1) return;
    // cur was flipped by the update kernel: the accepted residual is in rbuf[cur^1], the older in rbuf[cur]
    const double* rn = st->cur ? r0 : r1;
    const double* ro = st->cur ? r1 : r0;
    const double t1 = st->t1;
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = rn[i] + t1 * (rn[i] - ro[i]);
        a = fma(v, v, a);
    }
    a = block_sum(a, sh);
    if (threadIdx.x == 0) c0part[blockIdx.x] = a;
}

__global__ void fista_result_kernel(const FistaState* st, const double* beta, int64_t vstride, int d,
                                    double* __restrict__ w_out, const double* r0, const double* r1, int64_t n,
                                    double* __restrict__ r_out) {
    {   // batched mode: blockIdx.y = instance; outputs are [B][d] and [B][n]
        const int64_t g = blockIdx.y;
        st += g;
        beta += g * vstride;
        r0 += g * n;
        r1 += g * n;
        if (w_out) w_out += g * d;
        if (r_out) r_out += g * n;
    }
    const double* r = st->cur ? r1 : r0;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gs = (int64_t)gridDim.x * blockDim.x;
    if (w_out)
        for (int64_t c = gid; c < d; c += gs) w_out[c] = beta[c];
    if (r_out)
        for (int64_t i = gid; i < n; i += gs) r_out[i] = r[i];
}

// ---- margins m = D w - lambda / rho (algorithms.py:89) from a maintained Dw ----------------------
// `scal` (here and below, may be null): device block [rho, lam_fista, thr_f32] bound with rbl_bind_scalars — when
// present it overrides the by-value scalar so that a captured CUDA graph of the iteration can be replayed with
// new values
__global__ void margins_kernel(const double* __restrict__ Dw, const double* __restrict__ lam, double rho, int64_t n,
                               double* __restrict__ m, const double* __restrict__ scal) {
    rbl_pdl_wait();
    if (scal) rho = scal[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m[i] = Dw[i] - lam[i] / rho;
}

// EHRM clip of the isotonic prox at B (PAV_cpt.py:203-226,268-288): use_clip 1 = candidate 2, max(B, prox with
// sigma = b) (:213,271); use_clip 2 = candidate 1, min(B, prox with sigma = a) (:207,264); 0 = none.  With a bound
// scalar block (captured graphs) the mode travels in scal[3] so one graph serves both candidates.
__device__ __forceinline__ double apply_clip(double v, int use_clip, double clip) {
    if (use_clip == 1) return v < clip ? clip : v;
    if (use_clip == 2) return v > clip ? clip : v;
    return v;
}

// ---- EHRM: the two scalars the reference compares to pick its candidate (PAV_cpt.py:203-226) -----------------
// f1 = sum_i a_i log(1+e^{x1_i}) + rho/2 |x1 - m|^2 with x1 = min(prox_{sigma=a}(m), B),
// f2 = sum_i b_i log(1+e^{x2_i}) + rho/2 |x2 - m|^2 with x2 = max(prox_{sigma=b}(m), B)   (element level, sorted m).
// Fixed assignment of elements to threads and fixed-order reductions: bit-reproducible.
__global__ void ehrm_sums_kernel(const double* __restrict__ ms, const double* __restrict__ sa,
                                 const double* __restrict__ sb, double B, double rho, int64_t n,
                                 double* __restrict__ part, int np, const double* __restrict__ scal) {
    __shared__ double sh[33];
    if (scal) rho = scal[0];
    double l1 = 0.0, q1 = 0.0, l2 = 0.0, q2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double m = ms[i], a = sa[i], b = sb[i];
        double x1 = rbl_block_prox(RBL_LOSS_BCE, a, m, rho);
        double x2 = rbl_block_prox(RBL_LOSS_BCE, b, m, rho);
        if (x1 > B) x1 = B;    // :207
        if (x2 <= B) x2 = B;   // :213
        l1 = fma(a, rbl_log1pexp(x1), l1);
        q1 = fma(x1 - m, x1 - m, q1);
        l2 = fma(b, rbl_log1pexp(x2), l2);
        q2 = fma(x2 - m, x2 - m, q2);
    }
    l1 = block_sum(l1, sh);
    q1 = block_sum(q1, sh);
    l2 = block_sum(l2, sh);
    q2 = block_sum(q2, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = l1;
        part[np + blockIdx.x] = q1;
        part[2 * np + blockIdx.x] = l2;
        part[3 * np + blockIdx.x] = q2;
    }
}

__global__ void __launch_bounds__(1024) ehrm_sums_finalize_kernel(const double* __restrict__ part, int np,
                                                                  double rho, double* __restrict__ out2,
                                                                  const double* __restrict__ scal) {
    __shared__ double sh[33];
    if (scal) rho = scal[0];
    double v[4];
    for (int q = 0; q < 4; ++q) {
        double a = 0.0;
        for (int k = threadIdx.x; k < np; k += blockDim.x) a += part[q * np + k];
        v[q] = block_sum(a, sh);
    }
    if (threadIdx.x == 0) {
        out2[0] = v[0] + rho / 2 * v[1];
        out2[1] = v[2] + rho / 2 * v[3];
    }
}

// ---- z[perm] = z_sorted (algorithms.py:103-104), only the rows this rank owns; b = z + lambda/rho
__global__ void scatter_kernel(const double* __restrict__ zs, const int32_t* __restrict__ perm, int64_t n_global,
                               int64_t row_lo, int64_t n_local, int use_clip, double clip,
                               const double* __restrict__ lam, double rho, double* __restrict__ z,
                               double* __restrict__ b, const double* __restrict__ scal) {
    if (scal) {
        rho = scal[0];
        if (use_clip) use_clip = (int)scal[3];
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_global;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = (int64_t)perm[i] - row_lo;
        if (row >= 0 && row < n_local) {
            const double v = apply_clip(zs[i], use_clip, clip);
            z[row] = v;
            if (b) b[row] = v + lam[row] / rho;
        }
    }
}

// ---- scatter + the ACTIVE-ROW list -------------------------------------------------------------------------
// r0 = b - D w = z + lambda/rho - D w = z - m  is exactly zero wherever the prox is the identity: every rank
// with sigma_i = 0 outside the pooled blocks (80% of the rows for superquantile q = 0.8), and for the hinge
// every margin below the kink.  The gradient pass D^T r0 therefore only has to read the ACTIVE rows.  Two
// kernels: (1) scatter as before, with a blocked rank assignment, also counting the owned active ranks per
// CTA; (2) stable compaction in rank order (deterministic: every CTA re-derives its base from the counts).
constexpr int kActThreads = 256;

__device__ __forceinline__ bool active_item(const double* __restrict__ zs, const double* __restrict__ ms,
                                            const int32_t* __restrict__ perm, int64_t i, int64_t row_lo,
                                            int64_t n_local, int use_clip, double clip, int64_t* row_out,
                                            double* v_out, double* delta_out) {
    const int64_t row = (int64_t)perm[i] - row_lo;
    const double v = apply_clip(zs[i], use_clip, clip);
    *row_out = row;
    *v_out = v;
    const double dl = v - ms[i];
    *delta_out = dl;
    return row >= 0 && row < n_local && dl != 0.0;
}

__global__ void __launch_bounds__(kActThreads) scatter_active_kernel(
    const double* __restrict__ zs, const double* __restrict__ ms, const int32_t* __restrict__ perm, int64_t n_global,
    int64_t row_lo, int64_t n_local, int use_clip, double clip, const double* __restrict__ lam, double rho,
    double* __restrict__ z, double* __restrict__ b, int64_t chunk, int* __restrict__ cta_count,
    const double* __restrict__ scal) {
    rbl_pdl_wait();
    __shared__ int wcount[kActThreads / 32];
    if (scal) {
        rho = scal[0];
        if (use_clip) use_clip = (int)scal[3];
    }
    const int64_t r0 = (int64_t)blockIdx.x * chunk;
    const int64_t r1 = (r0 + chunk < n_global) ? r0 + chunk : n_global;
    int cnt = 0;
    for (int64_t i = r0 + threadIdx.x; i < r1; i += kActThreads) {
        int64_t row;
        double v, dl;
        const bool act = active_item(zs, ms, perm, i, row_lo, n_local, use_clip, clip, &row, &v, &dl);
        if (row >= 0 && row < n_local) {
            z[row] = v;
            if (b) b[row] = v + lam[row] / rho;
        }
        cnt += act ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) wcount[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kActThreads / 32; ++w) t += wcount[w];
        cta_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kActThreads) compact_active_kernel(
    const double* __restrict__ zs, const double* __restrict__ ms, const int32_t* __restrict__ perm, int64_t n_global,
    int64_t row_lo, int64_t n_local, int use_clip, double clip, int64_t chunk, const int* __restrict__ cta_count,
    int32_t* __restrict__ act_row, double* __restrict__ act_delta, int* __restrict__ act_total,
    const double* __restrict__ scal) {
    rbl_pdl_wait();
    if (scal && use_clip) use_clip = (int)scal[3];
    __shared__ int sh[kActThreads / 32 + 1];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // base = sum of the counts of the CTAs before me (and the grand total, written by the last CTA)
    int part = 0;
    for (int c = tid; c < (int)blockIdx.x; c += kActThreads) part += cta_count[c];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) sh[warp] = part;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kActThreads / 32; ++w) t += sh[w];
        s_base = t;
        if (blockIdx.x == gridDim.x - 1) *act_total = t + cta_count[blockIdx.x];
    }
    __syncthreads();
    int base = s_base;
    const int64_t r0 = (int64_t)blockIdx.x * chunk;
    const int64_t r1 = (r0 + chunk < n_global) ? r0 + chunk : n_global;
    for (int64_t i0 = r0; i0 < r1; i0 += kActThreads) {  // block-uniform trip count
        const int64_t i = i0 + tid;
        int64_t row = 0;
        double v, dl = 0.0;
        const bool act = (i < r1) && active_item(zs, ms, perm, i, row_lo, n_local, use_clip, clip, &row, &v, &dl);
        const unsigned m = __ballot_sync(0xffffffffu, act);
        __syncthreads();
        if (lane == 0) sh[warp] = __popc(m);
        __syncthreads();
        int off = base, tot = 0;
#pragma unroll
        for (int w = 0; w < kActThreads / 32; ++w) {
            const int t = sh[w];
            if (w < warp) off += t;
            tot += t;
        }
        if (act) {
            const int k = off + __popc(m & ((1u << lane) - 1u));
            act_row[k] = (int32_t)row;
            act_delta[k] = dl;
        }
        base += tot;
    }
}

// ---- dual update + primal residual (algorithms.py:132,135) --------------------------------------
// Dw is either given (from_residual = 0) or recovered from the last FISTA residual r = b - D w.
__global__ void dual_kernel(const double* __restrict__ z, double* __restrict__ Dw, const double* __restrict__ b,
                            const double* __restrict__ r, int from_residual, double* __restrict__ lam, double rho,
                            int64_t n, double* __restrict__ part) {
    __shared__ double sh[33];
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double dw;
        if (from_residual) {
            dw = b[i] - r[i];
            Dw[i] = dw;
        } else {
            dw = Dw[i];
        }
        const double res = z[i] - dw;
        lam[i] = __dadd_rn(lam[i], __dmul_rn(rho, res));  // two roundings, like numpy (:132)
        a = fma(res, res, a);
    }
    a = block_sum(a, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}

// out[0] = sum(part[0..np)), out[1] = ||w - w_prev||^2 (algorithms.py:136), out[2] = ||w||^2, out[3] = ||w||_1
__global__ void __launch_bounds__(1024) finalize_kernel(const double* __restrict__ part, int np,
                                                        const double* __restrict__ w,
                                                        const double* __restrict__ w_prev, int d,
                                                        double* __restrict__ out) {
    __shared__ double sh[33];
    double a = 0.0;
    for (int k = threadIdx.x; k < np; k += blockDim.x) a += part[k];
    a = block_sum(a, sh);
    double dd = 0.0, w2 = 0.0, w1 = 0.0;
    if (w) {
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            const double wc = w[c];
            if (w_prev) {
                const double df = wc - w_prev[c];
                dd = fma(df, df, dd);
            }
            w2 = fma(wc, wc, w2);
            w1 += fabs(wc);
        }
        dd = block_sum(dd, sh);
        w2 = block_sum(w2, sh);
        w1 = block_sum(w1, sh);
    }
    if (threadIdx.x == 0) {
        out[0] = a;
        out[1] = dd;
        out[2] = w2;
        out[3] = w1;
    }
}

// ---- support of w: ascending indices of the non-zero coordinates, compacted by a block scan (1 CTA) ------
__global__ void __launch_bounds__(1024) support_kernel(const double* __restrict__ w, int d, int32_t* __restrict__ idx,
                                                       double* __restrict__ val, int* __restrict__ nnz) {
    __shared__ int wsum[32];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < d; c0 += 1024) {
        const int c = c0 + tid;
        const double v = (c < d) ? w[c] : 0.0;
        const bool nz = (v != 0.0);
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int q = 0; q < warp; ++q) off += wsum[q];
        if (nz) {
            const int k = off + __popc(m & ((1u << lane) - 1u));
            idx[k] = c;
            val[k] = v;
        }
        __syncthreads();
        if (tid == 0) {
            int t = base;
            for (int q = 0; q < 32; ++q) t += wsum[q];
            base = t;
        }
        __syncthreads();
    }
    if (tid == 0) *nnz = base;
}

// ---- dual update with a SPARSE w (algorithms.py:132,135): Dw_i = sum_{k in supp} D[i][j_k] w_k ------------
// The l1 w-step leaves exact zeros in w; skipping them reads nnz 32-byte sectors per row instead of the
// whole 8 d-byte row (HBM traffic down by d / (4 nnz)).  A group of GS lanes owns a row (GS = 8, 16 or 32,
// the smallest that covers nnz), 4 rows in flight per group.  Runs only when nnz <= cap; the dense
// pass (RBL_PASS_DUAL) is gated the other way.
template <int GS, typename T>
__device__ __forceinline__ void sparse_dual_rows(const T* __restrict__ D, int64_t ld, int64_t n,
                                                 const int32_t* __restrict__ idx, const double* __restrict__ val,
                                                 int nnz, const double* __restrict__ z, double* __restrict__ Dw,
                                                 double* __restrict__ lam, double rho, double& acc) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % GS, grp = lane / GS;
    constexpr int GPW = 32 / GS;
    // warp-uniform loop bounds (the shuffles below need every lane of the warp): a warp owns GPW consecutive
    // rows per step, U steps in flight
    const int64_t wbase = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW;
    const int64_t gstride = (int64_t)gridDim.x * (blockDim.x >> 5) * GPW;
    constexpr int U = 4;  // rows in flight per group
    for (int64_t b0 = wbase; b0 < n; b0 += U * gstride) {
        const int64_t i0 = b0 + grp;
        double a[U], zv[U], lv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * gstride;
            a[u] = 0.0;
            zv[u] = 0.0;
            lv[u] = 0.0;
            if (sub == 0 && i < n) {  // issued now, consumed after the dot products
                zv[u] = z[i];
                lv[u] = lam[i];
            }
        }
        for (int k = sub; k < nnz; k += GS) {
            const int j = idx[k];
            const double wv = val[k];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * gstride;
                if (i < n) a[u] = fma((double)__ldg(&D[i * ld + j]), wv, a[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int o = GS / 2; o; o >>= 1) a[u] += __shfl_xor_sync(0xffffffffu, a[u], o);
            const int64_t i = i0 + u * gstride;
            if (sub == 0 && i < n) {
                const double res = zv[u] - a[u];
                Dw[i] = a[u];
                lam[i] = __dadd_rn(lv[u], __dmul_rn(rho, res));  // two roundings, like numpy (:132)
                acc = fma(res, res, acc);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) sparse_dual_kernel(const T* __restrict__ D, int64_t ld, int64_t n,
                                                          const int32_t* __restrict__ idx,
                                                          const double* __restrict__ val,
                                                          const int* __restrict__ nnz_ptr, int cap,
                                                          const double* __restrict__ z, double* __restrict__ Dw,
                                                          double* __restrict__ lam, double rho,
                                                          double* __restrict__ part,
                                                          const double* __restrict__ scal) {
    rbl_pdl_wait();
    __shared__ double sh[33];
    const int nnz = *nnz_ptr;
    if (nnz > cap) return;  // the dense pass handles it
    if (scal) rho = scal[0];
    double acc = 0.0;
    if (nnz <= 8) sparse_dual_rows<8>(D, ld, n, idx, val, nnz, z, Dw, lam, rho, acc);
    else if (nnz <= 16) sparse_dual_rows<16>(D, ld, n, idx, val, nnz, z, Dw, lam, rho, acc);
    else sparse_dual_rows<32>(D, ld, n, idx, val, nnz, z, Dw, lam, rho, acc);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// Same update from the TRANSPOSED copy Dt (d x n, row j = column j of D): the nnz touched columns are read as
// contiguous n-vectors, so the traffic is nnz n 8 bytes of fully coalesced loads (80 MB at nnz = 10, n = 1M)
// instead of one DRAM page activation per row of D (the sector gather above is activate-bound: ~2 TB/s).
template <typename T>
__global__ void __launch_bounds__(256) sparse_dual_t_kernel(const T* __restrict__ Dt, int64_t n,
                                                            const int32_t* __restrict__ idx,
                                                            const double* __restrict__ val,
                                                            const int* __restrict__ nnz_ptr, int cap,
                                                            const double* __restrict__ z, double* __restrict__ Dw,
                                                            double* __restrict__ lam, double rho,
                                                            double* __restrict__ part,
                                                            const double* __restrict__ scal) {
    rbl_pdl_wait();
    __shared__ double sh[33];
    __shared__ int s_idx[256];
    if (scal) rho = scal[0];
    __shared__ double s_val[256];
    const int nnz = *nnz_ptr;
    if (nnz > cap) return;  // the dense pass handles it
    double acc = 0.0;
    for (int k0 = 0; k0 == 0 || k0 < nnz; k0 += 256) {  // nnz <= cap <= d/16: one chunk in practice
        __syncthreads();
        if (k0 + (int)threadIdx.x < nnz) {
            s_idx[threadIdx.x] = idx[k0 + threadIdx.x];
            s_val[threadIdx.x] = val[k0 + threadIdx.x];
        }
        __syncthreads();
        const int kn = (nnz - k0 < 256) ? (nnz - k0) : 256;
        const bool first = (k0 == 0), last_chunk = (k0 + 256 >= nnz);
        for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n;
             i += (int64_t)gridDim.x * blockDim.x * 2) {
            using V2 = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
            const bool two = (i + 1 < n) && ((n & 1) == 0);  // two-element loads need aligned columns
            double a0 = first ? 0.0 : Dw[i], a1 = (two && !first) ? Dw[i + 1] : 0.0;
            if (two) {
#pragma unroll 4
                for (int k = 0; k < kn; ++k) {
                    const V2 v = __ldg(reinterpret_cast<const V2*>(Dt + (int64_t)s_idx[k] * n + i));
                    a0 = fma((double)v.x, s_val[k], a0);
                    a1 = fma((double)v.y, s_val[k], a1);
                }
            } else {
                for (int k = 0; k < kn; ++k) a0 = fma((double)__ldg(Dt + (int64_t)s_idx[k] * n + i), s_val[k], a0);
                if (i + 1 < n) {
                    a1 = first ? 0.0 : Dw[i + 1];
                    for (int k = 0; k < kn; ++k)
                        a1 = fma((double)__ldg(Dt + (int64_t)s_idx[k] * n + i + 1), s_val[k], a1);
                }
            }
            Dw[i] = a0;
            if (i + 1 < n) Dw[i + 1] = a1;
            if (last_chunk) {
                const double r0 = z[i] - a0;
                lam[i] = __dadd_rn(lam[i], __dmul_rn(rho, r0));  // two roundings, like numpy (:132)
                acc = fma(r0, r0, acc);
                if (i + 1 < n) {
                    const double r1 = z[i + 1] - a1;
                    lam[i + 1] = __dadd_rn(lam[i + 1], __dmul_rn(rho, r1));
                    acc = fma(r1, r1, acc);
                }
            }
        }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// Dt = D^T (d x n) through a 32 x 33 shared-memory tile
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ D, int64_t ld, int64_t n, int d,
                                                        T* __restrict__ Dt) {
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t ntr = (n + 31) / 32;
    const int ntc = (d + 31) / 32;
    for (int64_t t = blockIdx.x; t < ntr * ntc; t += gridDim.x) {
        const int64_t tr = t / ntc;
        const int tc = (int)(t - tr * ntc);
        const int64_t r0 = tr * 32;
        const int c0 = tc * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t r = r0 + ty + 8 * q;
            const int c = c0 + tx;
            tile[ty + 8 * q][tx] = (r < n && c < d) ? D[r * ld + c] : (T)0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + ty + 8 * q;
            const int64_t r = r0 + tx;
            if (c < d && r < n) Dt[(int64_t)c * n + r] = tile[tx][ty + 8 * q];
        }
        __syncthreads();
    }
}

// out8 = [||z - Dw||^2, ||w - w_prev||^2, ||w||^2, ||w||_1, nnz(w), 1 if the sparse kernel ran, k, sweeps]
__global__ void __launch_bounds__(1024) dual_finalize_kernel(const double* __restrict__ part_dense, int np_dense,
                                                             const double* __restrict__ part_sparse, int np_sparse,
                                                             const int* __restrict__ nnz_ptr, int cap,
                                                             const double* __restrict__ w,
                                                             const double* __restrict__ w_prev, int d,
                                                             const FistaState* __restrict__ st,
                                                             const int* __restrict__ act_total,
                                                             double* __restrict__ out,
                                                             double* __restrict__ w_copy) {
    rbl_pdl_wait();
    __shared__ double sh[33];
    const int nnz = *nnz_ptr;
    const bool sparse = nnz <= cap;
    const double* part = sparse ? part_sparse : part_dense;
    const int np = sparse ? np_sparse : np_dense;
    double a = 0.0;
    for (int k = threadIdx.x; k < np; k += blockDim.x) a += part[k];
    a = block_sum(a, sh);
    double dd = 0.0, w2 = 0.0, w1 = 0.0;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const double wc = w[c];
        if (w_copy) w_copy[c] = wc;  // e.g. pinned host memory: the iterate travels with the residuals
        if (w_prev) {
            const double df = wc - w_prev[c];
            dd = fma(df, df, dd);
        }
        w2 = fma(wc, wc, w2);
        w1 += fabs(wc);
    }
    dd = block_sum(dd, sh);
    w2 = block_sum(w2, sh);
    w1 = block_sum(w1, sh);
    if (threadIdx.x == 0) {
        out[0] = a;
        out[1] = dd;
        out[2] = w2;
        out[3] = w1;
        out[4] = (double)nnz;
        out[5] = sparse ? 1.0 : 0.0;
        out[6] = (double)st->k;       // last w-step: FISTA iterations, sweeps/passes (saves the host a poll)
        out[7] = (double)st->passes;
        out[8] = (double)*act_total;  // active rows of the last rbl_scatter_active
    }
}

// ---- objective: sum_i sigma_i * loss(u_(i)) over ascending margins (objective.py:71-81) ------------
// loss is non-decreasing in the margin u = D w, so sorting the margins sorts the losses.
__global__ void objective_kernel(const double* __restrict__ u_sorted, const double* __restrict__ sigma, int loss,
                                 int64_t n, double* __restrict__ part) {
    __shared__ double sh[33];
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        a = fma(sigma[i], rbl_margin_loss(loss, u_sorted[i]), a);
    a = block_sum(a, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}

}  // namespace

// ---- host launchers ----------------------------------------------------------------------------
int rbl_k_build_design(rbl_ctx* c, const double* X, int64_t ldx, const double* y, double* D, int64_t nrows,
                       cudaStream_t s) {
    if (c->esz == 4)
        build_design_kernel<float><<<c->num_sms * 8, 256, 0, s>>>(X, ldx, y, reinterpret_cast<float*>(D), c->ld, nrows,
                                                                c->d);
    else
        build_design_kernel<double><<<c->num_sms * 8, 256, 0, s>>>(X, ldx, y, D, c->ld, nrows, c->d);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_reduce_partials(rbl_ctx* c, int with_c0, const FistaState* st, cudaStream_t s) {
    const int threads = 256;
    const int blocks = (c->d + 31) / 32;
    RBL_CUDA(rbl_launch_pdl(reduce_partials_kernel, dim3(blocks), dim3(threads), 0, s, c->gpart, c->sspart, with_c0 ? c->c0part : nullptr,
                                                      c->pass_grid, c->vec_grid, c->ld, c->d, c->red, st));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_fista_update(rbl_ctx* c, cudaStream_t s) {
    fista_update_kernel<<<1, 1024, 0, s>>>(c->fista, c->d, c->red, 0, c->beta, c->beta_p, c->beta_prev, c->g_p,
                                           c->g_prev, 0, c->pow_tab);
    RBL_LAUNCH_CHECK();
    fista_combine_kernel<<<c->vec_grid, kVecThreads, 0, s>>>(c->fista, c->rbuf[0], c->rbuf[1], c->n_local, c->c0part);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_fista_result(rbl_ctx* c, double* w_out, double* r_out, cudaStream_t s) {
    fista_result_kernel<<<c->vec_grid, kVecThreads, 0, s>>>(c->fista, c->beta, 0, c->d, w_out, c->rbuf[0],
                                                           c->rbuf[1], c->n_local, r_out);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_margins(rbl_ctx* c, const double* Dw, const double* lam, double rho, double* m, cudaStream_t s) {
    RBL_CUDA(rbl_launch_pdl(margins_kernel, dim3(c->vec_grid), dim3(kVecThreads), 0, s, Dw, lam, rho, c->n_local, m, c->scal));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_scatter(rbl_ctx* c, const double* zs, const int32_t* perm, int use_clip, double clip, const double* lam,
                  double rho, double* z, double* b, cudaStream_t s) {
    scatter_kernel<<<c->vec_grid, kVecThreads, 0, s>>>(zs, perm, c->n_global, c->row_lo, c->n_local, use_clip, clip,
                                                      lam, rho, z, b, c->scal);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_ehrm_sums(rbl_ctx* c, const double* ms, const double* sa, const double* sb, double B, double rho,
                    double* out2, cudaStream_t s) {
    ehrm_sums_kernel<<<c->vec_grid, kVecThreads, 0, s>>>(ms, sa, sb, B, rho, c->n_global, c->vpart, c->vec_grid,
                                                        c->scal);
    RBL_LAUNCH_CHECK();
    ehrm_sums_finalize_kernel<<<1, 1024, 0, s>>>(c->vpart, c->vec_grid, rho, out2, c->scal);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_scatter_active(rbl_ctx* c, const double* zs, const double* ms, const int32_t* perm, int use_clip,
                         double clip, const double* lam, double rho, double* z, double* b, cudaStream_t s) {
    const int grid = c->vec_grid;
    const int64_t chunk = (c->n_global + grid - 1) / grid;
    RBL_CUDA(rbl_launch_pdl(scatter_active_kernel, dim3(grid), dim3(kActThreads), 0, s, zs, ms, perm, c->n_global, c->row_lo, c->n_local, use_clip,
                                                      clip, lam, rho, z, b, chunk, c->act_cta_count, c->scal));
    RBL_LAUNCH_CHECK();
    RBL_CUDA(rbl_launch_pdl(compact_active_kernel, dim3(grid), dim3(kActThreads), 0, s, zs, ms, perm, c->n_global, c->row_lo, c->n_local, use_clip,
                                                      clip, chunk, c->act_cta_count, c->act_row, c->act_delta,
                                                      c->act_total, c->scal));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_dual(rbl_ctx* c, const double* z, double* Dw, const double* b, const double* r, int from_residual,
               double* lam, double rho, const double* w, const double* w_prev, double* out4, cudaStream_t s) {
    dual_kernel<<<c->vec_grid, kVecThreads, 0, s>>>(z, Dw, b, r, from_residual, lam, rho, c->n_local, c->vpart);
    RBL_LAUNCH_CHECK();
    finalize_kernel<<<1, 1024, 0, s>>>(c->vpart, c->vec_grid, w, w_prev, c->d, out4);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

// support list, sparse kernel (gated nnz <= cap), then the caller launches the dense pass (gated nnz > cap)
int rbl_k_dual_sparse(rbl_ctx* c, const double* D, const double* Dt, const double* w, const double* z, double* Dw,
                      double* lam, double rho, int cap, int support_ready, cudaStream_t s) {
    if (!support_ready) {
        support_kernel<<<1, 1024, 0, s>>>(w, c->d, c->sup_idx, c->sup_val, c->sup_nnz);
        RBL_LAUNCH_CHECK();
    }
    if (Dt)
        if (c->esz == 4)
            RBL_CUDA(rbl_launch_pdl(sparse_dual_t_kernel<float>, dim3(c->vec_grid), dim3(kVecThreads), 0, s,
                                    reinterpret_cast<const float*>(Dt), c->n_local, c->sup_idx, c->sup_val, c->sup_nnz,
                                    cap, z, Dw, lam, rho, c->vpart, c->scal));
        else
            RBL_CUDA(rbl_launch_pdl(sparse_dual_t_kernel<double>, dim3(c->vec_grid), dim3(kVecThreads), 0, s, Dt,
                                    c->n_local, c->sup_idx, c->sup_val, c->sup_nnz, cap, z, Dw, lam, rho, c->vpart,
                                    c->scal));
    else if (c->esz == 4)
        RBL_CUDA(rbl_launch_pdl(sparse_dual_kernel<float>, dim3(c->vec_grid), dim3(kVecThreads), 0, s,
                                reinterpret_cast<const float*>(D), c->ld, c->n_local, c->sup_idx, c->sup_val,
                                c->sup_nnz, cap, z, Dw, lam, rho, c->vpart, c->scal));
    else
        RBL_CUDA(rbl_launch_pdl(sparse_dual_kernel<double>, dim3(c->vec_grid), dim3(kVecThreads), 0, s, D, c->ld,
                                c->n_local, c->sup_idx, c->sup_val, c->sup_nnz, cap, z, Dw, lam, rho, c->vpart,
                                c->scal));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_transpose(rbl_ctx* c, const double* D, double* Dt, cudaStream_t s) {
    if (c->esz == 4)
        transpose_kernel<float><<<c->num_sms * 16, 256, 0, s>>>(reinterpret_cast<const float*>(D), c->ld, c->n_local,
                                                                c->d, reinterpret_cast<float*>(Dt));
    else
        transpose_kernel<double><<<c->num_sms * 16, 256, 0, s>>>(D, c->ld, c->n_local, c->d, Dt);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_dual_finalize(rbl_ctx* c, int cap, const double* w, const double* w_prev, double* out8, double* w_copy,
                        cudaStream_t s) {
    RBL_CUDA(rbl_launch_pdl(dual_finalize_kernel, dim3(1), dim3(1024), 0, s, c->sspart, c->pass_grid, c->vpart, c->vec_grid, c->sup_nnz, cap, w,
                                            w_prev, c->d, c->fista, c->act_total, out8, w_copy));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_finalize(rbl_ctx* c, const double* part, int np, const double* w, const double* w_prev, double* out4,
                   cudaStream_t s) {
    finalize_kernel<<<1, 1024, 0, s>>>(part, np, w, w_prev, c->d, out4);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_objective(rbl_ctx* c, const double* u_sorted, const double* sigma, int loss, const double* w, double* out4,
                    cudaStream_t s) {
    objective_kernel<<<c->vec_grid, kVecThreads, 0, s>>>(u_sorted, sigma, loss, c->n_global, c->vpart);
    RBL_LAUNCH_CHECK();
    finalize_kernel<<<1, 1024, 0, s>>>(c->vpart, c->vec_grid, w, nullptr, c->d, out4);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

// ---- batched (K10) launchers: B instances, buffers laid out [B][...] --------------------------------
int rbl_k_fista_update_batch(rbl_ctx* c, int g0, int ng, cudaStream_t s) {
    const int64_t vs = c->ld + 8;
    fista_update_kernel<<<ng, 1024, 0, s>>>(c->bfista + g0, c->d, c->bred + (size_t)g0 * vs, vs,
                                            c->bbeta + (size_t)g0 * vs, c->bbeta_p + (size_t)g0 * vs,
                                            c->bbeta_prev + (size_t)g0 * vs, c->bg_p + (size_t)g0 * vs,
                                            c->bg_prev + (size_t)g0 * vs, vs, c->pow_tab);
    RBL_LAUNCH_CHECK();
    dim3 grid(c->vec_grid, ng);
    fista_combine_kernel<<<grid, kVecThreads, 0, s>>>(c->bfista + g0, c->brbuf[0] + (size_t)g0 * c->n_local,
                                                     c->brbuf[1] + (size_t)g0 * c->n_local, c->n_local,
                                                     c->bc0part + (size_t)g0 * c->vec_grid);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_fista_result_batch(rbl_ctx* c, int B, double* w_out, double* r_out, cudaStream_t s) {
    dim3 grid(c->vec_grid, B);
    fista_result_kernel<<<grid, kVecThreads, 0, s>>>(c->bfista, c->bbeta, c->ld + 8, c->d, w_out, c->brbuf[0],
                                                    c->brbuf[1], c->n_local, r_out);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}
