// pav_kernels.cu — rank-based z-step prox on the sorted margins: element prox (Newton / closed form),
// then the weighted pool-adjacent-violators (isotonic) projection as a balanced tree of merges.
//
// Replaces: PAV_solver.__init__/get_opt (src/util/pav.py:54-178), individual_solver
// (src/util/individual_solver.py:90-130) and, for EHRM, PAV_solver_CPT (src/util/PAV_cpt.py:169-293;
// as shipped == max(B, isotonic prox with sigma = b); the clip is applied by the scatter kernel).
// Algorithm and its derivation: csrc/pav_core.h (shared with the CPU emulation in tests/native/).
//
// Kernels
//   chunk_prefix   : chunk-local exclusive prefix sums in double-double (sigma once, margins per call)
//   chunk_offsets  : exclusive scan of the chunk totals (one small CTA)
//   pav_chunk      : one CTA per 1024-element chunk; prox per element, local prefix of m, then the
//                    10 in-chunk merge levels entirely in shared memory
//   pav_level      : one CTA per pair of solved ranges for the remaining log2(n/1024) levels; one
//                    thread runs the O(log n) merge search, the whole CTA fills the pooled block
#include "common.cuh"
#include "pav_core.h"

namespace {

constexpr int kChunkLog2 = 10;
constexpr int kChunk = 1 << kChunkLog2;
constexpr int kPavThreads = 256;
constexpr int kPer = kChunk / kPavThreads;  // 4 consecutive elements per thread

__device__ __forceinline__ dd_t shfl_up_dd(dd_t v, int o) {
    dd_t r;
    r.hi = __shfl_up_sync(0xffffffffu, v.hi, o);
    r.lo = __shfl_up_sync(0xffffffffu, v.lo, o);
    return r;
}

// exclusive scan of one dd value per thread across the block (256 threads); returns the exclusive
// prefix for this thread and the block total in *total.  sh: >= 2*8 doubles of shared memory.
__device__ __forceinline__ dd_t block_excl_scan_dd(dd_t v, dd_t* total, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kPavThreads / 32;
    dd_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        dd_t y = shfl_up_dd(x, o);
        if (lane >= o) x = dd_add(y, x);
    }
    __syncthreads();
    if (lane == 31) {
        sh[2 * warp] = x.hi;
        sh[2 * warp + 1] = x.lo;
    }
    __syncthreads();
    dd_t woff = dd_make(0.0), tot = dd_make(0.0);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        dd_t t;
        t.hi = sh[2 * w];
        t.lo = sh[2 * w + 1];
        if (w < warp) woff = dd_add(woff, t);
        tot = dd_add(tot, t);
    }
    *total = tot;
    // exclusive = warp offset + (inclusive - own)
    dd_t incl_prev = shfl_up_dd(x, 1);
    if (lane == 0) incl_prev = dd_make(0.0);
    return dd_add(woff, incl_prev);
}

// chunk-local exclusive prefix (dd) of x[0..n): loc[i] for every i, loc[n] if the last chunk is partial,
// and the chunk total
__global__ void __launch_bounds__(kPavThreads) chunk_prefix_kernel(const double* __restrict__ x, int64_t n,
                                                                   double* __restrict__ loc_hi,
                                                                   double* __restrict__ loc_lo,
                                                                   double* __restrict__ tot_hi,
                                                                   double* __restrict__ tot_lo) {
    __shared__ double sh[16];
    const int64_t base = (int64_t)blockIdx.x * kChunk;
    const int64_t i0 = base + (int64_t)threadIdx.x * kPer;
    double v[kPer];
    dd_t run = dd_make(0.0);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        v[q] = (i0 + q < n) ? x[i0 + q] : 0.0;
        run = dd_add_d(run, v[q]);
    }
    dd_t total;
    dd_t ex = block_excl_scan_dd(run, &total, sh);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        if (i0 + q < n) {
            loc_hi[i0 + q] = ex.hi;
            loc_lo[i0 + q] = ex.lo;
        }
        ex = dd_add_d(ex, v[q]);
    }
    if (threadIdx.x == 0) {
        tot_hi[blockIdx.x] = total.hi;
        tot_lo[blockIdx.x] = total.lo;
        if (n - base < kChunk) {  // entry n closes a partial last chunk
            loc_hi[n] = total.hi;
            loc_lo[n] = total.lo;
        }
    }
}

// off[c] = sum_{c' < c} tot[c'], c = 0..nch   (single CTA)
__global__ void __launch_bounds__(kPavThreads) chunk_offsets_kernel(const double* __restrict__ tot_hi,
                                                                    const double* __restrict__ tot_lo, int64_t nch,
                                                                    double* __restrict__ off_hi,
                                                                    double* __restrict__ off_lo) {
    __shared__ double sh[16];
    const int64_t per = (nch + kPavThreads - 1) / kPavThreads;
    const int64_t c0 = (int64_t)threadIdx.x * per;
    dd_t run = dd_make(0.0);
    for (int64_t c = c0; c < c0 + per && c < nch; ++c) {
        dd_t t;
        t.hi = tot_hi[c];
        t.lo = tot_lo[c];
        run = dd_add(run, t);
    }
    dd_t total;
    dd_t ex = block_excl_scan_dd(run, &total, sh);
    for (int64_t c = c0; c < c0 + per && c < nch; ++c) {
        off_hi[c] = ex.hi;
        off_lo[c] = ex.lo;
        dd_t t;
        t.hi = tot_hi[c];
        t.lo = tot_lo[c];
        ex = dd_add(ex, t);
    }
    if (threadIdx.x == 0) {
        off_hi[nch] = total.hi;
        off_lo[nch] = total.lo;
    }
}

struct ChunkSmem {
    double val[kChunk];
    double psh[kChunk + 1], psl[kChunk + 1];
    double pmh[kChunk + 1], pml[kChunk + 1];
    double sh[16];
    int64_t rec_lo[16], rec_hi[16];
    double rec_v[16];
};

__global__ void __launch_bounds__(kPavThreads) pav_chunk_kernel(
    int loss, double rho, const double* __restrict__ sigma, const double* __restrict__ m, int64_t n,
    const double* __restrict__ ps_loc_hi, const double* __restrict__ ps_loc_lo, const double* __restrict__ ps_tot_hi,
    const double* __restrict__ ps_tot_lo, double* __restrict__ pm_loc_hi, double* __restrict__ pm_loc_lo,
    double* __restrict__ pm_tot_hi, double* __restrict__ pm_tot_lo, double* __restrict__ val_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChunkSmem& S = *reinterpret_cast<ChunkSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kChunk;
    const int len = (int)((n - base < kChunk) ? (n - base) : kChunk);
    const int l0 = tid * kPer;

    // element prox + local prefix of m (exclusive, dd) ; sigma's local prefix was built once at setup
    double mv[kPer];
    dd_t run = dd_make(0.0);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int li = l0 + q;
        if (li < len) {
            const double sg = sigma[base + li];
            mv[q] = m[base + li];
            S.val[li] = rbl_block_prox(loss, sg, mv[q], rho);
            S.psh[li] = ps_loc_hi[base + li];
            S.psl[li] = ps_loc_lo[base + li];
        } else {
            mv[q] = 0.0;
        }
        run = dd_add_d(run, mv[q]);
    }
    dd_t total;
    dd_t ex = block_excl_scan_dd(run, &total, S.sh);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int li = l0 + q;
        if (li < len) {
            S.pmh[li] = ex.hi;
            S.pml[li] = ex.lo;
            pm_loc_hi[base + li] = ex.hi;
            pm_loc_lo[base + li] = ex.lo;
        }
        ex = dd_add_d(ex, mv[q]);
    }
    if (tid == 0) {
        pm_tot_hi[blockIdx.x] = total.hi;
        pm_tot_lo[blockIdx.x] = total.lo;
        // closing entries of the exclusive prefixes
        S.pmh[len] = total.hi;
        S.pml[len] = total.lo;
        if (len < kChunk) {
            pm_loc_hi[base + len] = total.hi;
            pm_loc_lo[base + len] = total.lo;
        }
        S.psh[len] = (len == kChunk) ? ps_tot_hi[blockIdx.x] : ps_loc_hi[base + len];
        S.psl[len] = (len == kChunk) ? ps_tot_lo[blockIdx.x] : ps_loc_lo[base + len];
    }
    __syncthreads();

    PrefixFlat ps{S.psh, S.psl}, pm{S.pmh, S.pml};
    for (int w = 1; w < len; w <<= 1) {
        const int npairs = (len + 2 * w - 1) / (2 * w);
        if (2 * w <= 32) {
            // many small pairs: the searching thread fills its own pooled block
            for (int pr = tid; pr < npairs; pr += kPavThreads) {
                const int a = pr * 2 * w, b = a + w;
                if (b < len) {
                    const int c = (a + 2 * w < len) ? a + 2 * w : len;
                    int64_t lo, hi;
                    double v;
                    if (pav_merge_search(loss, rho, S.val, ps, pm, a, b, c, &lo, &hi, &v))
                        for (int64_t i = lo; i < hi; ++i) S.val[i] = v;
                }
            }
            __syncthreads();
        } else {
            // few large pairs (<= 16): search by one thread each, fill by the whole CTA
            if (tid < npairs) {
                const int a = tid * 2 * w, b = a + w;
                int64_t lo = 0, hi = 0;
                double v = 0.0;
                if (b < len) {
                    const int c = (a + 2 * w < len) ? a + 2 * w : len;
                    if (!pav_merge_search(loss, rho, S.val, ps, pm, a, b, c, &lo, &hi, &v)) lo = hi = 0;
                }
                S.rec_lo[tid] = lo;
                S.rec_hi[tid] = hi;
                S.rec_v[tid] = v;
            }
            __syncthreads();
            for (int i = tid; i < len; i += kPavThreads) {
                const int pr = i / (2 * w);
                if (i >= S.rec_lo[pr] && i < S.rec_hi[pr]) S.val[i] = S.rec_v[pr];
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < len; i += kPavThreads) val_out[base + i] = S.val[i];
}

// element-wise prox without pooling (PAV level 0 as a standalone op; individual_solver.py:112-130)
__global__ void prox_elementwise_kernel(int loss, double rho, const double* __restrict__ sigma,
                                        const double* __restrict__ m, int64_t n, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = rbl_block_prox(loss, sigma[i], m[i], rho);
}

// one CTA per pair of solved ranges of width w (w >= chunk)
__global__ void __launch_bounds__(kPavThreads) pav_level_kernel(int loss, double rho, double* __restrict__ val,
                                                                int64_t n, int64_t w, PrefixChunked ps,
                                                                PrefixChunked pm) {
    __shared__ int64_t s_lo, s_hi;
    __shared__ double s_v;
    const int64_t a = (int64_t)blockIdx.x * 2 * w, b = a + w;
    if (b >= n) return;
    const int64_t c = (a + 2 * w < n) ? a + 2 * w : n;
    if (threadIdx.x == 0) {
        int64_t lo = 0, hi = 0;
        double v = 0.0;
        if (!pav_merge_search(loss, rho, val, ps, pm, a, b, c, &lo, &hi, &v)) lo = hi = 0;
        s_lo = lo;
        s_hi = hi;
        s_v = v;
    }
    __syncthreads();
    const int64_t lo = s_lo, hi = s_hi;
    const double v = s_v;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kPavThreads) val[i] = v;
}

}  // namespace

int rbl_pav_chunk_log2() { return kChunkLog2; }

int rbl_k_prox_elementwise(rbl_ctx* c, int loss, const double* sigma, const double* m, int64_t n, double rho,
                           double* out, cudaStream_t s) {
    prox_elementwise_kernel<<<c->vec_grid, 256, 0, s>>>(loss, rho, sigma, m, n, out);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_prefix(rbl_ctx* c, const double* x, int64_t n, double* loc_hi, double* loc_lo, double* tot_hi,
                 double* tot_lo, double* off_hi, double* off_lo, cudaStream_t s) {
    const int64_t nch = (n + kChunk - 1) / kChunk;
    (void)c;
    chunk_prefix_kernel<<<(unsigned)nch, kPavThreads, 0, s>>>(x, n, loc_hi, loc_lo, tot_hi, tot_lo);
    RBL_LAUNCH_CHECK();
    chunk_offsets_kernel<<<1, kPavThreads, 0, s>>>(tot_hi, tot_lo, nch, off_hi, off_lo);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

// sigma: rank-order weights; ps_tot = per-chunk totals of sigma kept from setup
int rbl_k_pav(rbl_ctx* c, int loss, const double* m_sorted, double rho, double* z_sorted, cudaStream_t s) {
    const int64_t n = c->n_global;
    const int64_t nch = c->nchunks;
    static bool attr_set = false;
    if (!attr_set) {
        RBL_CUDA(cudaFuncSetAttribute(pav_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(ChunkSmem)));
        attr_set = true;
    }
    pav_chunk_kernel<<<(unsigned)nch, kPavThreads, sizeof(ChunkSmem), s>>>(
        loss, rho, c->sigma, m_sorted, n, c->ps_loc_hi, c->ps_loc_lo, c->ps_tot_hi, c->ps_tot_lo, c->pm_loc_hi,
        c->pm_loc_lo, c->ch_tot_hi, c->ch_tot_lo, z_sorted);
    RBL_LAUNCH_CHECK();
    if (nch > 1) {
        chunk_offsets_kernel<<<1, kPavThreads, 0, s>>>(c->ch_tot_hi, c->ch_tot_lo, nch, c->pm_off_hi, c->pm_off_lo);
        RBL_LAUNCH_CHECK();
        PrefixChunked ps{c->ps_loc_hi, c->ps_loc_lo, c->ps_off_hi, c->ps_off_lo, kChunkLog2};
        PrefixChunked pm{c->pm_loc_hi, c->pm_loc_lo, c->pm_off_hi, c->pm_off_lo, kChunkLog2};
        for (int64_t w = kChunk; w < n; w <<= 1) {
            const int64_t npairs = (n + 2 * w - 1) / (2 * w);
            pav_level_kernel<<<(unsigned)npairs, kPavThreads, 0, s>>>(loss, rho, z_sorted, n, w, ps, pm);
            RBL_LAUNCH_CHECK();
        }
    }
    return RBL_OK;
}
