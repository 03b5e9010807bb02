// gram_kernels.cu — the w-step on the Gram matrix G = D^T D (d x d) instead of on D (n x d).
//
// The reference precomputes DTD = D.T @ D in Optimizer.__init__ (src/optim/algorithms.py:24) and uses it
// in the l2 gradient (src/util/w_LBFGS.py:39-45).  Here it carries the WHOLE inner loop of the w-step:
// the LASSO / ridge sub-problem  min_w 1/2 ||b - D w||^2 + g(w)  is a quadratic in w, so with
//     w0  = the warm start (the previous ADMM iterate),
//     g0  = D^T (b - D w0),  ss0 = ||b - D w0||^2         (ONE fused pass over D per ADMM iteration)
// every quantity FISTA (src/util/fast_lasso.py:40-67) or L-BFGS-B (w_LBFGS.py:31-45) asks for follows from
// d x d products with G:
//     D^T (b - D beta)            = g0 - G (beta - w0)
//     ||b - D beta||^2            = ss0 - 2 (beta - w0).g0 + (beta - w0).G (beta - w0)
//     LHS - RHS of the line search (fast_lasso.py:50-56)
//        = [ ||b - D beta||^2 - ||b - D beta_p||^2 ] - [ L ||D||^2 - 2 D.g_p ]       (D := beta - beta_p)
//        = D.G D - L ||D||^2                          (exactly: the -2 D.g_p terms cancel)
// so a line-search trial costs one read of G (8 MB at d = 1000, L2-resident on B200: 126 MB L2) instead of
// one read of D (8 GB at n = 1M): an ADMM iteration needs exactly TWO passes over D (D^T b here, D w for the
// dual update) however many trials FISTA takes.  Working in the displacement beta - w0 keeps the
// cancellation in g0 - G(.) at the level of the step, not of ||G|| ||w||.
//
// Kernels
//   gram_step_kernel<2> : v = G (beta - w0), u = G (beta - beta_p) in one sweep over G (a warp per row,
//                         128-bit loads, both right-hand sides staged in shared memory), then the LAST CTA
//                         to finish (ticket) runs the FISTA control flow for that trial — accept/reject,
//                         momentum, stop test, next trial point — so a trial is ONE launch and the host
//                         never synchronises inside the inner loop.
//   gram_step_kernel<1> : q = G (w - w0) and, in the last CTA, red = [g0 - q, ss0 - 2 dw.g0 + dw.q]: the
//                         f/g evaluation L-BFGS-B asks for (w_LBFGS.py:31-45), same layout as rbl_fused_pass.
//   gram_syrk_kernel    : G = D^T D over this rank's rows on the FP64 tensor-core path (DMMA m8n8k4);
//                         algorithms.py:24.  Row-sharded jobs all-reduce G once at construction.
// All reductions run in a fixed order: results are bit-reproducible and identical on every rank.
#include "common.cuh"

namespace {

constexpr int kGThreads = 256;
constexpr int kGWarps = kGThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// sum over the block (kGThreads threads), valid in every thread; sh holds >= kGWarps + 1 doubles
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kGWarps; ++w) t += sh[w];
        sh[kGWarps] = t;
    }
    __syncthreads();
    return sh[kGWarps];
}

struct GramParams {
    const double* G;
    int64_t ldg;
    int d;
    FistaState* st;
    const double* w0;     // warm start (beta0)
    const double* red0;   // [g0 (d), ss0]
    double *beta, *beta_p, *beta_prev, *q_p, *q_prev, *g_p;
    double* xs;           // [2][ldg] right-hand sides: beta - w0, beta - beta_p
    double* vu;           // [2][ldg] products
    unsigned int* ticket;
    const float* pow_tab;
    // eval mode (NRHS = 1)
    const double* w;      // evaluation point
    double* red_out;      // [d + 2]
};

// ---- FISTA control flow for one trial (fast_lasso.py:44-65), run by one CTA of kGThreads threads ------
// init: seeds the iteration at beta_p = beta_prev = w0 (fast_lasso.py:32-39) and forms the first trial.
__device__ void gram_fista_update(const GramParams& p, bool init, double* sh) {
    FistaState* st = p.st;
    const int tid = threadIdx.x, nt = kGThreads, d = p.d;
    const FistaState S = *st;
    __syncthreads();  // everyone holds S before thread 0 rewrites *st
    const double* g0 = p.red0;
    double* xs0 = p.xs;
    double* xs1 = p.xs + p.ldg;
    float L_prev = S.L_prev, L_cur = S.L_cur;
    int i_k = S.i_k, k = S.k, done = 0;
    double t = S.t, t1 = S.t1, crit = S.crit, lhs = 0.0;

    if (init) {
        for (int c = tid; c < d; c += nt) {
            const double bw = p.w0[c];
            p.beta_p[c] = bw;
            p.beta_prev[c] = bw;
            p.q_p[c] = 0.0;
            p.q_prev[c] = 0.0;
            p.g_p[c] = g0[c];
        }
        k = 0;
        i_k = 0;
        t = 1.0;
        L_cur = __fmul_rn(L_prev, p.pow_tab[0]);
    } else {
        // LHS > RHS  <=>  D.G D > L ||D||^2 with D = beta - beta_p (header); false for NaN like the reference
        double a = 0.0;
        for (int c = tid; c < d; c += nt) a = fma(xs1[c], __ldcg(&p.vu[p.ldg + c]), a);
        lhs = block_sum(a, sh);
        if (lhs > S.rhs) {
            ++i_k;  // :45-46
            L_cur = __fmul_rn(L_prev, p.pow_tab[i_k < 127 ? i_k : 127]);
        } else {
            L_prev = L_cur;                                              // :58
            const double tnext = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;  // :59
            t1 = (t - 1.0) / tnext;                                      // :61
            double a2 = 0.0;
            for (int c = tid; c < d; c += nt) {
                const double df = p.beta[c] - p.beta_prev[c];  // :60
                a2 = fma(df, df, a2);
            }
            crit = sqrt(block_sum(a2, sh));  // :63
            ++k;
            if (crit < S.tol || k >= S.max_iter) {
                done = 1;  // result is beta
            } else {
                t = tnext;
                for (int c = tid; c < d; c += nt) {
                    const double bc = p.beta[c];
                    const double v = __ldcg(&p.vu[c]);         // G (beta - w0)
                    const double df = bc - p.beta_prev[c];
                    p.beta_p[c] = bc + t1 * df;                // :62
                    const double qp = v + t1 * (v - p.q_prev[c]);  // G (beta_p - w0): affine in beta
                    p.q_p[c] = qp;
                    p.q_prev[c] = v;
                    p.g_p[c] = g0[c] - qp;                     // D^T (b - D beta_p), :41-43
                    p.beta_prev[c] = bc;
                }
                i_k = 0;
                L_cur = __fmul_rn(L_prev, p.pow_tab[0]);
            }
        }
    }
    double rhs = S.rhs;
    if (!done) {
        __syncthreads();
        // trial: beta = soft(beta_p + g_p / L_cur, lam / L_cur)   (:47-49)
        const double Ld = (double)L_cur;
        const double thr = S.thr_f32 ? (double)__fdiv_rn((float)S.lam, L_cur) : S.lam / Ld;
        double r1 = 0.0;
        for (int c = tid; c < d; c += nt) {
            const double bp = p.beta_p[c], g = p.g_p[c];
            const double bs = bp + g / Ld;
            const double mag = fmax(fabs(bs) - thr, 0.0);
            const double sgn = (bs > 0.0) ? 1.0 : ((bs < 0.0) ? -1.0 : 0.0);
            const double bn = mag * sgn;
            p.beta[c] = bn;
            const double df = bn - bp;
            xs0[c] = bn - p.w0[c];
            xs1[c] = df;
            r1 = fma(df, df, r1);
        }
        r1 = block_sum(r1, sh);
        rhs = Ld * r1;  // L ||D||^2 (:51-53 without the common -2 D.g_p term)
    }
    if (tid == 0) {
        st->t = t;
        st->rhs = rhs;
        st->t1 = t1;
        st->crit = crit;
        st->ss_last = lhs;
        st->L_prev = L_prev;
        st->L_cur = L_cur;
        st->i_k = i_k;
        st->k = k;
        st->done = done;
        st->passes = S.passes + (init ? 0 : 1);  // sweeps over G
        st->trials = S.trials + (init ? 0 : 1);
    }
}

__global__ void __launch_bounds__(kGThreads) gram_fista_init_kernel(const GramParams p) {
    __shared__ double sh[kGWarps + 1];
    gram_fista_update(p, true, sh);
}

// NRHS = 2: FISTA trial; NRHS = 1: f/g evaluation at p.w
template <int NRHS>
__global__ void __launch_bounds__(kGThreads) gram_step_kernel(const GramParams p) {
    extern __shared__ __align__(16) double xsm[];  // [NRHS][ldg]
    __shared__ double sh[kGWarps + 1];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = p.d;
    const int64_t ldg = p.ldg;
    if (NRHS == 2 && p.st->done) return;  // converged earlier in this batch of enqueued steps
    if (NRHS == 2) {
        for (int c = tid; c < ldg; c += kGThreads) {
            xsm[c] = (c < d) ? p.xs[c] : 0.0;
            xsm[ldg + c] = (c < d) ? p.xs[ldg + c] : 0.0;
        }
    } else {
        for (int c = tid; c < ldg; c += kGThreads) xsm[c] = (c < d) ? p.w[c] - p.w0[c] : 0.0;
    }
    __syncthreads();
    const int row = blockIdx.x * kGWarps + warp;
    if (row < d) {
        const double2* g2 = reinterpret_cast<const double2*>(p.G + (size_t)row * ldg);
        const double2* x0 = reinterpret_cast<const double2*>(xsm);
        const double2* x1 = reinterpret_cast<const double2*>(xsm + ldg);
        const int ld2 = (int)(ldg >> 1);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll 4
        for (int c2 = lane; c2 < ld2; c2 += 32) {
            const double2 g = __ldg(&g2[c2]);
            const double2 xa = x0[c2];
            a0 = fma(g.x, xa.x, a0);
            a1 = fma(g.y, xa.y, a1);
            if (NRHS == 2) {
                const double2 xb = x1[c2];
                b0 = fma(g.x, xb.x, b0);
                b1 = fma(g.y, xb.y, b1);
            }
        }
        const double a = warp_sum(a0 + a1);
        if (NRHS == 2) {
            const double b = warp_sum(b0 + b1);
            if (lane == 0) {
                p.vu[row] = a;
                p.vu[ldg + row] = b;
            }
        } else if (lane == 0) {
            p.vu[row] = a;
        }
    }
    // ---- last CTA to arrive consumes the products
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(p.ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) *p.ticket = 0u;  // self-cleaning for the next launch
    __threadfence();
    if (NRHS == 2) {
        gram_fista_update(p, false, sh);
    } else {
        // red_out = [D^T (b - D w) (d), ||b - D w||^2]  (header identities)
        const double* g0 = p.red0;
        double a = 0.0, bq = 0.0;
        for (int c = tid; c < d; c += kGThreads) {
            const double q = __ldcg(&p.vu[c]);
            const double dw = xsm[c];
            p.red_out[c] = g0[c] - q;
            a = fma(dw, g0[c], a);
            bq = fma(dw, q, bq);
        }
        a = block_sum(a, sh);
        bq = block_sum(bq, sh);
        if (tid == 0) {
            p.red_out[d] = g0[d] - 2.0 * a + bq;
            p.red_out[d + 1] = 0.0;
        }
    }
}

// ---- the whole FISTA call as ONE persistent cooperative kernel ------------------------------------------
// CTA c owns rows [c rpc, (c+1) rpc) of G (kept in shared memory when they fit: 8 MB of G spread over 148 SMs is
// 56 KB each at d = 1000) and a private copy of the d-vector state in shared memory.  Per trial: every CTA
// multiplies its rows by the two right-hand sides, publishes the 2 rpc products, ONE grid barrier, then every
// CTA redundantly runs the same control flow on the same numbers (bit-identical by construction), so no
// second barrier and no host round trip is needed until the call has converged.
struct GramPersist {
    const double* G;
    int64_t ldg;
    int d;
    FistaState* st;       // final state for the host (k, sweeps, L, crit)
    const double* w0;
    const double* red0;
    double* vu;           // [2 (parity)][2][ldg]
    double* beta_out;     // global copy of the final iterate (handle-internal)
    double* w_out;        // caller's w (may be null; may alias w0: it is written after the last grid barrier)
    double* w_prev_out;   // receives w0 (may be null)
    int32_t* sup_idx;     // ascending support of the result (may be null): indices, values, count
    double* sup_val;
    int* sup_nnz;
    unsigned int* bar;    // [0] barrier counter, [1] exit ticket; both zero on entry and on exit
    const float* pow_tab;
    int rpc;
    int g_in_smem;
    double lam, tol;
    float L0;
    int thr_f32, max_iter;
    const double* scal;   // device block [rho, lam, thr_f32] overriding lam / thr_f32 when bound (may be null)
    unsigned long long* dbg;  // dev tool: %globaltimer stamps of CTA 0 (null: off)
};

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// sums of KC per-thread partials over the block, results valid in every thread (fixed order: deterministic and
// identical in every CTA); sh holds KC * kGWarps doubles
template <int KC>
__device__ __forceinline__ void block_sum_vec(double (&v)[KC], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < KC; ++c) v[c] = warp_sum(v[c]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < KC; ++c) sh[c * kGWarps + warp] = v[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kGWarps; ++w) t += sh[c * kGWarps + w];
        v[c] = t;
    }
}

// KC line-search candidates per sweep.  The reference restarts L at 17 on every call and walks
// L = 17 * 2.5^i up to ~lambda_max(G) ~ 1e6: ~13 rejected trials before the first accepted one.  All trial
// points of one FISTA iteration depend only on (beta_p, g_p, L_i), so KC of them are formed at once, multiplied
// by G in ONE sweep (KC right-hand sides Delta_i = beta_i - beta_p; G (beta_i - w0) = q_p + G Delta_i) and the
// first i that passes the test is taken — the same accept/reject sequence as trying them one by one, at one
// grid barrier per KC trials.
template <int KC>
__global__ void __launch_bounds__(kGThreads, 1) gram_fista_persistent_kernel(const GramPersist p) {
    extern __shared__ __align__(16) double psm[];
    __shared__ double sh[KC * kGWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = kGThreads;
    const int d = p.d;
    const int64_t ld = p.ldg;
    double* W0 = psm;
    double* G0 = W0 + ld;
    double* BETA = G0 + ld;
    double* BETA_P = BETA + ld;
    double* BETA_PREV = BETA_P + ld;
    double* Q_P = BETA_PREV + ld;
    double* Q_PREV = Q_P + ld;
    double* G_P = Q_PREV + ld;
    double* DEL = G_P + ld;             // [KC][ld] candidate steps beta_i - beta_p
    double* Gs = DEL + (size_t)KC * ld;  // [rpc][ld] when g_in_smem
    const int row0 = blockIdx.x * p.rpc;
    const int row1 = (row0 + p.rpc < d) ? row0 + p.rpc : d;

    for (int c = tid; c < ld; c += nt) {
        const double bw = (c < d) ? p.w0[c] : 0.0, g = (c < d) ? p.red0[c] : 0.0;
        W0[c] = bw;
        G0[c] = g;
        BETA[c] = bw;
        BETA_P[c] = bw;
        BETA_PREV[c] = bw;
        Q_P[c] = 0.0;
        Q_PREV[c] = 0.0;
        G_P[c] = g;
#pragma unroll
        for (int q = 0; q < KC; ++q) DEL[(size_t)q * ld + c] = 0.0;  // padding columns stay zero
    }
    if (p.g_in_smem) {
        const int ld2 = (int)(ld >> 1);
        const double2* src = reinterpret_cast<const double2*>(p.G + (size_t)row0 * ld);
        double2* dst = reinterpret_cast<double2*>(Gs);
        const int total = (row1 - row0) * ld2;
        for (int e = tid; e < total; e += nt) dst[e] = __ldg(&src[e]);
    }
    const double lam_ = p.scal ? p.scal[1] : p.lam;
    const int thr_f32_ = p.scal ? (p.scal[2] != 0.0 ? 1 : 0) : p.thr_f32;
    __shared__ float s_pow[128];
    for (int i = tid; i < 128; i += nt) s_pow[i] = p.pow_tab[i];
    float L_prev = p.L0, L_acc = p.L0;
    int i_k0 = 0, i_k = 0, k = 0, sweeps = 0, trials = 0, par = 0;
    // candidates evaluated per sweep: all KC, always.  (While L climbs from L0 the first iteration of a call rejects
    // ~12 candidates; once one has been accepted L_prev is right and i = 0 passes almost always — the later sweeps could
    // be narrowed, but a sweep's cost is its barrier and reduction latency, not the arithmetic of the candidates.)
    int kc = KC;
    double t = 1.0, t1 = 0.0, crit = 0.0, lhs_acc = 0.0, rhs_acc = 0.0;
    unsigned int target = 0;
    __syncthreads();
    int dbg_n = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 60) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
            p.dbg[1 + dbg_n++] = tt;
        }
    };
    stamp();

    while (true) {
        // ---- KC candidates: beta_i = soft(beta_p + g_p / L_i, lam / L_i), L_i = L_prev * eta^(i_k0 + i)
        float Lc[KC];
        double Ld[KC], thr[KC], r1[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            const int ii = i_k0 + q;
            Lc[q] = __fmul_rn(L_prev, s_pow[ii < 127 ? ii : 127]);  // fast_lasso.py:46
            Ld[q] = (double)Lc[q];
            thr[q] = thr_f32_ ? (double)__fdiv_rn((float)lam_, Lc[q]) : lam_ / Ld[q];
            r1[q] = 0.0;
        }
        for (int c = tid; c < d; c += nt) {
            const double bp = BETA_P[c], g = G_P[c];
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                if (q >= kc) break;
                const double bs = bp + g / Ld[q];  // :47
                const double mag = fmax(fabs(bs) - thr[q], 0.0);
                const double sgn = (bs > 0.0) ? 1.0 : ((bs < 0.0) ? -1.0 : 0.0);
                const double df = mag * sgn - bp;  // :49-50
                DEL[(size_t)q * ld + c] = df;
                r1[q] = fma(df, df, r1[q]);
            }
        }
        block_sum_vec<KC>(r1, sh);  // its barriers also publish DEL
        stamp();
        // ---- u_i = G Delta_i for my rows
        double* vu = p.vu + (size_t)par * KC * ld;
        for (int row = row0 + warp; row < row1; row += kGWarps) {
            const double2* g2 = p.g_in_smem ? reinterpret_cast<const double2*>(Gs + (size_t)(row - row0) * ld)
                                            : reinterpret_cast<const double2*>(p.G + (size_t)row * ld);
            const int ld2 = (int)(ld >> 1);
            double ax[KC], ay[KC];
#pragma unroll
            for (int q = 0; q < KC; ++q) ax[q] = ay[q] = 0.0;
            for (int c2 = lane; c2 < ld2; c2 += 32) {
                const double2 g = p.g_in_smem ? g2[c2] : __ldg(&g2[c2]);
#pragma unroll
                for (int q = 0; q < KC; ++q) {
                    if (q >= kc) break;
                    const double2 x = reinterpret_cast<const double2*>(DEL + (size_t)q * ld)[c2];
                    ax[q] = fma(g.x, x.x, ax[q]);
                    ay[q] = fma(g.y, x.y, ay[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                if (q >= kc) break;
                const double a = warp_sum(ax[q] + ay[q]);
                if (lane == 0) __stcg(&vu[(size_t)q * ld + row], a);
            }
        }
        ++sweeps;
        stamp();
        target += gridDim.x;
        grid_barrier(p.bar, target);
        stamp();
        // ---- line search, identical in every CTA: first i with NOT (Delta_i.G Delta_i > L_i ||Delta_i||^2)
        double lhs[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) lhs[q] = 0.0;
        for (int c = tid; c < d; c += nt) {
#pragma unroll
            for (int q = 0; q < KC; ++q) lhs[q] = fma(DEL[(size_t)q * ld + c], __ldcg(&vu[(size_t)q * ld + c]), lhs[q]);
        }
        block_sum_vec<KC>(lhs, sh);
        stamp();
        int acc_q = -1;
#pragma unroll
        for (int q = 0; q < KC; ++q)
            if (acc_q < 0 && !(lhs[q] > Ld[q] * r1[q])) acc_q = q;  // cond = (LHS > RHS), false for NaN (:56)
        par ^= 1;
        if (acc_q < 0) {  // all KC rejected: next KC candidates from the same (beta_p, g_p)
            i_k0 += KC;
            trials += KC;
            __syncthreads();
            continue;
        }
        i_k = i_k0 + acc_q;
        trials += acc_q + 1;
        L_acc = Lc[acc_q];
        lhs_acc = lhs[acc_q];
        rhs_acc = Ld[acc_q] * r1[acc_q];
        const double Lda = Ld[acc_q], thra = thr[acc_q];
        const double* ua = vu + (size_t)acc_q * ld;
        L_prev = L_acc;                                              // :58
        const double tnext = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;  // :59
        t1 = (t - 1.0) / tnext;                                      // :61
        double a2[1] = {0.0};
        for (int c = tid; c < d; c += nt) {
            const double bp = BETA_P[c], g = G_P[c];
            const double bs = bp + g / Lda;  // the accepted trial point, same expression as above
            const double mag = fmax(fabs(bs) - thra, 0.0);
            const double sgn = (bs > 0.0) ? 1.0 : ((bs < 0.0) ? -1.0 : 0.0);
            const double bn = mag * sgn;
            BETA[c] = bn;
            const double df = bn - BETA_PREV[c];  // :60
            a2[0] = fma(df, df, a2[0]);
        }
        block_sum_vec<1>(a2, sh);
        crit = sqrt(a2[0]);  // :63
        ++k;
        if (crit < p.tol || k >= p.max_iter) break;
        t = tnext;
        for (int c = tid; c < d; c += nt) {
            const double bc = BETA[c];
            const double v = Q_P[c] + __ldcg(&ua[c]);      // G (beta - w0) = G (beta_p - w0) + G Delta
            const double df = bc - BETA_PREV[c];
            BETA_P[c] = bc + t1 * df;                      // :62
            const double qp = v + t1 * (v - Q_PREV[c]);    // G (beta_p - w0): affine in beta
            Q_P[c] = qp;
            Q_PREV[c] = v;
            G_P[c] = G0[c] - qp;                           // D^T (b - D beta_p), :41-43
            BETA_PREV[c] = bc;
        }
        i_k0 = 0;
        __syncthreads();
        stamp();
    }
    stamp();
    if (blockIdx.x == 0) {
        for (int c = tid; c < d; c += nt) {
            const double b = BETA[c];
            p.beta_out[c] = b;
            if (p.w_out) p.w_out[c] = b;
            if (p.w_prev_out) p.w_prev_out[c] = W0[c];
        }
        if (p.sup_idx) {  // ascending support of beta for the sparse dual pass (saves the support kernel)
            __shared__ int s_wc[kGWarps];
            int base = 0;
            for (int c0 = 0; c0 < d; c0 += nt) {
                const int c = c0 + tid;
                const double v = (c < d) ? BETA[c] : 0.0;
                const bool nz = (v != 0.0);
                const unsigned m = __ballot_sync(0xffffffffu, nz);
                __syncthreads();
                if (lane == 0) s_wc[warp] = __popc(m);
                __syncthreads();
                int off = base, tot = 0;
#pragma unroll
                for (int w = 0; w < kGWarps; ++w) {
                    if (w < warp) off += s_wc[w];
                    tot += s_wc[w];
                }
                if (nz) {
                    const int k2 = off + __popc(m & ((1u << lane) - 1u));
                    p.sup_idx[k2] = c;
                    p.sup_val[k2] = v;
                }
                base += tot;
            }
            if (tid == 0) *p.sup_nnz = base;
        }
        if (tid == 0) {
            FistaState* st = p.st;
            st->t = t;
            st->rhs = rhs_acc;
            st->t1 = t1;
            st->crit = crit;
            st->ss_last = lhs_acc;
            st->tol = p.tol;
            st->lam = lam_;
            st->L_prev = L_prev;
            st->L_cur = L_acc;
            st->i_k = i_k;
            st->k = k;
            st->max_iter = p.max_iter;
            st->done = 1;
            st->passes = sweeps;   // sweeps over G (each carries KC candidates)
            st->trials = trials;   // line-search trials consumed, as the reference would count them
            st->thr_f32 = thr_f32_;
        }
    }
    // leave the barrier words clean for the next launch: the last CTA out resets them (no memset node needed)
    __syncthreads();
    stamp();
    if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[0] = (unsigned long long)dbg_n;
    if (tid == 0) {
        const unsigned int tk = atomicAdd(p.bar + 1, 1u);
        if (tk == gridDim.x - 1) {
            p.bar[0] = 0u;
            p.bar[1] = 0u;
            __threadfence();
        }
    }
}

__global__ void gram_result_kernel(const double* __restrict__ beta, int d, double* __restrict__ w_out) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d; c += gridDim.x * blockDim.x) w_out[c] = beta[c];
}

// ---- G = D^T D over the local rows (algorithms.py:24) on the FP64 tensor cores --------------------------
// CTA tile: 64 x 64 of G, K-chunks of 32 rows of D staged (transposed access is free: A = D^T is read
// "row.col" straight from the row-major tile).  Split-K over row slabs; partials are summed in a fixed
// order by gram_syrk_reduce_kernel.  Only tiles with tj <= ti are computed, the reduce mirrors them.
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kST = 64;    // tile edge of G
constexpr int kSK = 16;    // rows of D per stage
constexpr int kSLd = 68;   // padded smem leading dimension (doubles): 4 mod 16 -> conflict-free fragment loads

struct SyrkParams {
    const void* D;
    int64_t ld;
    int64_t n;
    int d;
    int ntile;       // tiles per edge
    int nslab;       // split-K factor
    double* part;    // [nslab][npairs][64*64]
};

__device__ __forceinline__ void pair_to_tiles(int pair, int* ti, int* tj) {
    int t = 0;
    while (pair > t) {  // pair index -> (ti, tj) with tj <= ti
        pair -= t + 1;
        ++t;
    }
    *ti = t;
    *tj = pair;
}

template <typename T>
__global__ void __launch_bounds__(256) gram_syrk_kernel(const SyrkParams p) {
    const T* __restrict__ Dp = reinterpret_cast<const T*>(p.D);
    __shared__ __align__(16) double As[2][kSK][kSLd];  // rows k, columns of tile ti
    __shared__ __align__(16) double Bs[2][kSK][kSLd];  // rows k, columns of tile tj
    int ti, tj;
    pair_to_tiles(blockIdx.x, &ti, &tj);
    const int slab = blockIdx.y;
    const int64_t rows_per_slab = ((p.n + p.nslab - 1) / p.nslab + kSK - 1) / kSK * kSK;
    const int64_t r_begin = (int64_t)slab * rows_per_slab;
    const int64_t r_end = (r_begin + rows_per_slab < p.n) ? r_begin + rows_per_slab : p.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    // warp tile: 32 (i) x 16 (j): wi = (warp & 1) * 32, wj = (warp >> 1) * 16
    const int wi = (warp & 1) * 32, wj = (warp >> 1) * 16;
    double acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int ci0 = ti * kST, cj0 = tj * kST;
    // stage = 16 rows x 64 columns per operand = 1024 doubles / 256 threads = 4 each
    double ra[4], rb[4];
    auto fetch = [&](int64_t r0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;
            const int r = e >> 6, c = e & 63;
            const int64_t gr = r0 + r;
            const bool ok = gr < r_end;
            ra[q] = (ok && ci0 + c < p.d) ? (double)__ldg(&Dp[gr * p.ld + ci0 + c]) : 0.0;
            rb[q] = (ok && cj0 + c < p.d) ? (double)__ldg(&Dp[gr * p.ld + cj0 + c]) : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;
            As[buf][e >> 6][e & 63] = ra[q];
            Bs[buf][e >> 6][e & 63] = rb[q];
        }
    };
    int buf = 0;
    if (r_begin < r_end) {
        fetch(r_begin);
        stash(0);
    }
    __syncthreads();
    for (int64_t r0 = r_begin; r0 < r_end; r0 += kSK) {
        const bool more = r0 + kSK < r_end;
        if (more) fetch(r0 + kSK);  // global loads in flight during the MMAs
#pragma unroll
        for (int kk = 0; kk < kSK; kk += 4) {
            double af[4], bf[2];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[buf][kk + fk][wi + a * 8 + fr];  // A[m = i][k] = D[k][i]
#pragma unroll
            for (int b = 0; b < 2; ++b) bf[b] = Bs[buf][kk + fk][wj + b * 8 + fr];  // B[k][n = j] = D[k][j]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    double* out = p.part + ((size_t)slab * gridDim.x + blockIdx.x) * (kST * kST);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = wi + a * 8 + fr, j = wj + b * 8 + 2 * fk;  // C[m = fr][n = 2 fk, 2 fk + 1]
            out[i * kST + j] = acc[a][b][0];
            out[i * kST + j + 1] = acc[a][b][1];
        }
}

// fixed-order sum of the split-K partials; the lower triangle is mirrored so G is exactly symmetric
__global__ void gram_syrk_reduce_kernel(const double* __restrict__ part, int nslab, int npairs, int d, int64_t ldg,
                                        double* __restrict__ G, int accumulate) {
    int ti, tj;
    pair_to_tiles(blockIdx.x, &ti, &tj);
    for (int e = threadIdx.x; e < kST * kST; e += blockDim.x) {
        const int li = e >> 6, lj = e & 63;
        if (ti == tj && lj > li) continue;
        double s = 0.0;
        for (int k = 0; k < nslab; ++k) s += part[((size_t)k * npairs + blockIdx.x) * (kST * kST) + e];
        const int i = ti * kST + li, j = tj * kST + lj;
        if (i < d && j < d) {
            if (accumulate) s += G[(size_t)i * ldg + j];  // chunked build: same value on both sides of the diagonal
            G[(size_t)i * ldg + j] = s;
            G[(size_t)j * ldg + i] = s;
        }
    }
}


// ---- small-problem l1 w-step: scikit-learn's Lasso coordinate descent on G (algorithms.py:194-197) ----------------
// The reference solves n <= 500, d <= 60 problems with sklearn.linear_model.Lasso(alpha, tol=1e-8,
// fit_intercept=False, max_iter=50000): cyclic coordinate descent from w = 0 on  1/2 ||b - D w||^2 + l1 ||w||_1
// (l1 = alpha * n), residual kept up to date, and after every sweep whose largest coordinate change is below
// tol * max|w_j| the duality gap against tol * b.b decides (scikit-learn 1.2.2, _cd_fast.pyx
// enet_coordinate_descent; restated on the CPU in oracle/pav_oracle.c).  Here the same recursion runs on
// q = D^T b and G = D^T D: the vector c = D^T (b - D w) = q - G w plays the residual's part (tmp_j = c_j + G_jj w_j,
// c -= (w_j' - w_j) G_j), so a sweep costs d^2 flops and no pass over D.  ONE warp: lane l holds c_l and c_{l+32};
// G sits in shared memory.  q and b.b come from the warm-start pass [g0 = D^T (b - D w_ref), ||b - D w_ref||^2].
__global__ void __launch_bounds__(32) lasso_cd_gram_kernel(const double* __restrict__ G, int64_t ldg, int d,
                                                           const double* __restrict__ w_ref,
                                                           const double* __restrict__ red0, double l1, double tol,
                                                           int max_iter, double* __restrict__ w_out,
                                                           double* __restrict__ info) {
    __shared__ double Gs[64 * 64];
    __shared__ double ws[64], qs[64];
    const int lane = threadIdx.x;
    const unsigned full = 0xffffffffu;
    for (int i = lane; i < 64 * 64; i += 32) {
        const int r = i >> 6, k = i & 63;
        Gs[i] = (r < d && k < d) ? G[(int64_t)r * ldg + k] : 0.0;
    }
    for (int k = lane; k < 64; k += 32) ws[k] = k < d ? w_ref[k] : 0.0;
    __syncwarp();
    // q = g0 + G w_ref,  yy = b.b = ss0 + 2 w_ref.g0 + w_ref.(G w_ref)
    double yy_part = 0.0;
    for (int j = lane; j < 64; j += 32) {
        double gw = 0.0;
        for (int k = 0; k < d; ++k) gw = fma(Gs[j * 64 + k], ws[k], gw);
        const double g0 = j < d ? red0[j] : 0.0;
        qs[j] = g0 + gw;
        yy_part += ws[j] * (2.0 * g0 + gw);
    }
    for (int o = 16; o; o >>= 1) yy_part += __shfl_xor_sync(full, yy_part, o);
    const double yy = red0[d] + yy_part;
    __syncwarp();
    double c0 = qs[lane], c1 = qs[lane + 32];  // c = q - G w with w = 0
    double w0 = 0.0, w1 = 0.0;                 // lane l owns w_l and w_{l+32}
    const double gap_tol = tol * yy;
    double gap = tol + 1.0;
    int it = 0;
    for (it = 0; it < max_iter; ++it) {
        double w_max = 0.0, d_w_max = 0.0;
        for (int j = 0; j < d; ++j) {
            const double gjj = Gs[j * 64 + j];
            if (gjj == 0.0) continue;
            const int src = j & 31;
            const double cj = __shfl_sync(full, j < 32 ? c0 : c1, src);
            const double wj = __shfl_sync(full, j < 32 ? w0 : w1, src);
            const double tmp = cj + gjj * wj;
            const double a = fabs(tmp) - l1;
            const double wn = (tmp > 0.0 ? 1.0 : (tmp < 0.0 ? -1.0 : 0.0)) * (a > 0.0 ? a : 0.0) / gjj;
            const double dl = wn - wj;
            if (dl != 0.0) {
                c0 = fma(-dl, Gs[j * 64 + lane], c0);
                c1 = fma(-dl, Gs[j * 64 + lane + 32], c1);
                if (lane == src) {
                    if (j < 32) w0 = wn; else w1 = wn;
                }
            }
            d_w_max = fmax(d_w_max, fabs(dl));
            w_max = fmax(w_max, fabs(wn));
        }
        if (w_max == 0.0 || d_w_max / w_max < tol || it == max_iter - 1) {
            // duality gap (formulation A): R.R = yy - w.q - w.c,  R.y = yy - w.q,  X^T R = c
            double dn = fmax(fabs(c0), fabs(c1));
            double wq = w0 * qs[lane] + w1 * qs[lane + 32];
            double wc = w0 * c0 + w1 * c1;
            double l1n = fabs(w0) + fabs(w1);
            for (int o = 16; o; o >>= 1) {
                dn = fmax(dn, __shfl_xor_sync(full, dn, o));
                wq += __shfl_xor_sync(full, wq, o);
                wc += __shfl_xor_sync(full, wc, o);
                l1n += __shfl_xor_sync(full, l1n, o);
            }
            const double R2 = yy - wq - wc, Ry = yy - wq;
            double cst;
            if (dn > l1) {
                cst = l1 / dn;
                gap = 0.5 * (R2 + R2 * cst * cst);
            } else {
                cst = 1.0;
                gap = R2;
            }
            gap += l1 * l1n - cst * Ry;
            if (gap < gap_tol) {
                ++it;
                break;
            }
        }
    }
    if (lane < d) w_out[lane] = w0;
    if (lane + 32 < d) w_out[lane + 32] = w1;
    if (lane == 0 && info) {
        info[0] = (double)it;
        info[1] = gap;
        info[2] = gap_tol;
    }
}

}  // namespace

// ---- host launchers ------------------------------------------------------------------------------------
static GramParams gram_params(rbl_ctx* c, const double* G, const double* w0, const double* red0) {
    GramParams p;
    p.G = G;
    p.ldg = c->ld;
    p.d = c->d;
    p.st = c->fista;
    p.w0 = w0;
    p.red0 = red0;
    p.beta = c->beta;
    p.beta_p = c->beta_p;
    p.beta_prev = c->beta_prev;
    p.q_p = c->g_prev;  // the stream-mode g_prev slot is free in Gram mode
    p.q_prev = c->gq_prev;
    p.g_p = c->g_p;
    p.xs = c->gxs;
    p.vu = c->gvu;
    p.ticket = c->gticket;
    p.pow_tab = c->pow_tab;
    p.w = nullptr;
    p.red_out = nullptr;
    return p;
}

static int gram_set_smem(rbl_ctx* c, size_t smem) {
    RBL_PER_DEVICE(size_t, attr2, c);
    RBL_PER_DEVICE(size_t, attr1, c);
    if (smem > 48 * 1024) {
        if (smem > attr2) {
            RBL_CUDA(cudaFuncSetAttribute(gram_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr2 = smem;
        }
        if (smem > attr1) {
            RBL_CUDA(cudaFuncSetAttribute(gram_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr1 = smem;
        }
    }
    return RBL_OK;
}

int rbl_k_gram_fista_init(rbl_ctx* c, const double* G, const double* w0, const double* red0, cudaStream_t s) {
    c->gram_w0 = w0;
    c->gram_red0 = red0;
    const GramParams p = gram_params(c, G, w0, red0);
    gram_fista_init_kernel<<<1, kGThreads, 0, s>>>(p);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_gram_fista_steps(rbl_ctx* c, const double* G, int nsteps, cudaStream_t s) {
    const GramParams p = gram_params(c, G, c->gram_w0, c->gram_red0);
    const size_t smem = 2 * (size_t)c->ld * sizeof(double);
    int rc = gram_set_smem(c, smem);
    if (rc != RBL_OK) return rc;
    const int grid = (c->d + kGWarps - 1) / kGWarps;
    for (int i = 0; i < nsteps; ++i) {
        gram_step_kernel<2><<<grid, kGThreads, smem, s>>>(p);
        RBL_LAUNCH_CHECK();
    }
    return RBL_OK;
}

int rbl_k_gram_result(rbl_ctx* c, double* w_out, cudaStream_t s) {
    gram_result_kernel<<<(c->d + 255) / 256, 256, 0, s>>>(c->beta, c->d, w_out);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_gram_eval(rbl_ctx* c, const double* G, const double* w0, const double* red0, const double* w,
                    double* red_out, cudaStream_t s) {
    GramParams p = gram_params(c, G, w0, red0);
    p.w = w;
    p.red_out = red_out;
    const size_t smem = (size_t)c->ld * sizeof(double);
    int rc = gram_set_smem(c, 2 * smem);
    if (rc != RBL_OK) return rc;
    const int grid = (c->d + kGWarps - 1) / kGWarps;
    gram_step_kernel<1><<<grid, kGThreads, smem, s>>>(p);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

// scratch doubles rbl_k_gram_build needs for its split-K partials
static void syrk_shape(rbl_ctx* c, int64_t nrows, int* ntile, int* npairs, int* nslab) {
    *ntile = (c->d + kST - 1) / kST;
    *npairs = *ntile * (*ntile + 1) / 2;
    // enough CTAs to fill the machine ~4x over, at least 256 rows per slab
    int64_t want = ((int64_t)c->num_sms * 4 + *npairs - 1) / *npairs;
    const int64_t max_slabs = (nrows + 255) / 256;
    if (want > max_slabs) want = max_slabs;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    *nslab = (int)want;
}

size_t rbl_gram_scratch_doubles(rbl_ctx* c, int64_t nrows) {
    int ntile, npairs, nslab;
    syrk_shape(c, nrows, &ntile, &npairs, &nslab);
    return (size_t)nslab * npairs * kST * kST;
}

// G (+)= D_rows^T D_rows over `nrows` rows starting at D (accumulate != 0 adds to the G already there)
int rbl_k_gram_build(rbl_ctx* c, const double* D, int64_t nrows, int accumulate, double* G, double* scratch,
                     cudaStream_t s) {
    SyrkParams p;
    int npairs;
    syrk_shape(c, nrows, &p.ntile, &npairs, &p.nslab);
    p.D = D;
    p.ld = c->ld;
    p.n = nrows;
    p.d = c->d;
    p.part = scratch;
    if (!accumulate) RBL_CUDA(cudaMemsetAsync(G, 0, (size_t)c->d * c->ld * sizeof(double), s));
    dim3 grid(npairs, p.nslab);
    if (c->esz == 4) gram_syrk_kernel<float><<<grid, 256, 0, s>>>(p);
    else gram_syrk_kernel<double><<<grid, 256, 0, s>>>(p);
    RBL_LAUNCH_CHECK();
    gram_syrk_reduce_kernel<<<npairs, 256, 0, s>>>(scratch, p.nslab, npairs, c->d, c->ld, G, accumulate);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

// ---- persistent FISTA: shape + launch ----------------------------------------------------------------------
template <int KC>
static bool gp_try(rbl_ctx* c, int dev_max, int grid, int rpc) {
    const size_t vec_bytes = (size_t)(8 + KC) * c->ld * sizeof(double);
    if (vec_bytes + 2048 > (size_t)dev_max) return false;
    const size_t g_bytes = (size_t)rpc * c->ld * sizeof(double);
    const int g_in = (vec_bytes + g_bytes + 2048 <= (size_t)dev_max) ? 1 : 0;
    const size_t smem = vec_bytes + (g_in ? g_bytes : 0);
    if (cudaFuncSetAttribute(gram_fista_persistent_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gram_fista_persistent_kernel<KC>, kGThreads, smem) !=
            cudaSuccess || per_sm * c->num_sms < grid) {
        cudaGetLastError();
        return false;
    }
    c->gp_kc = KC;
    c->gp_g_in_smem = g_in;
    c->gp_smem = smem;
    return true;
}

// returns 0 if the persistent kernel cannot hold its state in shared memory (caller uses the per-trial launches)
int rbl_gram_persist_config(rbl_ctx* c) {
    if (c->gp_checked) return c->gp_grid > 0;
    c->gp_checked = 1;
    c->gp_grid = 0;
    int dev_max = 0, coop = 0;
    if (cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device) != cudaSuccess || !coop) return 0;
    int grid = c->num_sms;
    if (grid > c->d) grid = c->d;
    const int rpc = (c->d + grid - 1) / grid;
    grid = (c->d + rpc - 1) / rpc;
    // as many candidates per sweep as fit next to the state; G rows in shared memory if they fit as well
    if (!(gp_try<8>(c, dev_max, grid, rpc) || gp_try<4>(c, dev_max, grid, rpc) || gp_try<2>(c, dev_max, grid, rpc) ||
          gp_try<1>(c, dev_max, grid, rpc)))
        return 0;
    c->gp_rpc = rpc;
    c->gp_grid = grid;
    return 1;
}

int rbl_k_gram_fista_run(rbl_ctx* c, const double* G, const double* w0, const double* red0, double lam, int thr_f32,
                         float L0, double tol, int max_iter, double* w_out, double* w_prev_out, int with_support,
                         cudaStream_t s) {
    GramPersist p;
    p.w_prev_out = w_prev_out;
    p.sup_idx = with_support ? c->sup_idx : nullptr;
    p.sup_val = c->sup_val;
    p.sup_nnz = c->sup_nnz;
    p.G = G;
    p.ldg = c->ld;
    p.d = c->d;
    p.st = c->fista;
    p.w0 = w0;
    p.red0 = red0;
    p.vu = c->gvu2;
    p.beta_out = c->beta;
    p.w_out = w_out;
    p.bar = c->gticket + 8;
    p.pow_tab = c->pow_tab;
    p.rpc = c->gp_rpc;
    p.g_in_smem = c->gp_g_in_smem;
    p.lam = lam;
    p.tol = tol;
    p.L0 = L0;
    p.thr_f32 = thr_f32;
    p.max_iter = max_iter;
    p.scal = c->scal;
    p.dbg = c->sort_dbg;  // shared dev-tool buffer (rbl_sort_debug)
    void* args[] = {(void*)&p};
    const void* fn = c->gp_kc == 8   ? (const void*)gram_fista_persistent_kernel<8>
                     : c->gp_kc == 4 ? (const void*)gram_fista_persistent_kernel<4>
                     : c->gp_kc == 2 ? (const void*)gram_fista_persistent_kernel<2>
                                     : (const void*)gram_fista_persistent_kernel<1>;
    RBL_CUDA(cudaLaunchCooperativeKernel(fn, dim3(c->gp_grid), dim3(kGThreads), args, c->gp_smem, s));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_lasso_cd_gram(rbl_ctx* c, const double* G, const double* w_ref, const double* red0, double l1, double tol,
                        int max_iter, double* w_out, double* info, cudaStream_t s) {
    lasso_cd_gram_kernel<<<1, 32, 0, s>>>(G, c->ld, c->d, w_ref, red0, l1, tol, max_iter, w_out, info);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}
