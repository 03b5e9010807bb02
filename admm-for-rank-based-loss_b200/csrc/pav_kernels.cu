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
//   pav_chunk      : one CTA per 1024-element chunk: element prox + the 10 in-chunk merge levels in
//                    shared memory
//   pav_tree       : the remaining log2(n/1024) levels in ONE launch: one CTA per pair of chunks climbs
//                    the merge tree ("last arriver continues": no CTA ever waits for another)
//
// Few-segment path.  The element prox z(sigma, m) is non-decreasing in m and non-increasing in sigma, so on
// sorted margins it is ALREADY isotonic over every maximal run of ranks where sigma does not increase:
// violations can only start where the spectrum steps UP.  ERM has no such step (no pooling ever),
// superquantile has two (0 -> fractional weight -> 1/(n(1-q))), AoRR at most two.  When the spectrum has
// <= kMaxSeg such runs (found once in rbl_set_spectrum) the level-0 solved ranges are those runs, and the
// whole tree collapses to (#runs - 1) merges done by one warp over a lazily-overlaid value array, followed by
// a grid-wide fill of the (few) pooled blocks: ~0.1 ms at n = 1M however large the pooled block is.
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
                                                                   double* __restrict__ tot_lo,
                                                                   const double* __restrict__ sigma, int loss,
                                                                   double rho, double* __restrict__ prox_out,
                                                                   const double* __restrict__ scal) {
    rbl_pdl_wait();
    __shared__ double sh[16];
    if (scal) rho = scal[0];
    // chunks are taken from the TOP ranks down: rank-based spectra put their weight (and so the Newton solves of the
    // element prox) on the largest margins — those CTAs start first instead of forming the tail of the grid
    const int64_t chunk = (int64_t)gridDim.x - 1 - blockIdx.x;
    const int64_t base = chunk * kChunk;
    const int64_t i0 = base + (int64_t)threadIdx.x * kPer;
    double v[kPer];
    dd_t run = dd_make(0.0);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        v[q] = (i0 + q < n) ? x[i0 + q] : 0.0;
        run = dd_add_d(run, v[q]);
    }
    if (prox_out) {  // few-segment path: the element prox rides along with the prefix of the margins
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (i0 + q < n) prox_out[i0 + q] = rbl_block_prox(loss, sigma[i0 + q], v[q], rho);
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
        tot_hi[chunk] = total.hi;
        tot_lo[chunk] = total.lo;
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
    int64_t rec_lo[16], rec_hi[16];
    double rec_v[16];
};

struct TreeParams {
    int loss;
    double rho;
    const double* scal;  // device scalar block (rho first) overriding `rho` when bound; may be null
    const double* sigma;
    const double* m;
    int64_t n;
    const double *ps_loc_hi, *ps_loc_lo, *ps_tot_hi, *ps_tot_lo, *ps_off_hi, *ps_off_lo;
    const double *pm_loc_hi, *pm_loc_lo, *pm_tot_hi, *pm_tot_lo, *pm_off_hi, *pm_off_lo;
    double* val;
    unsigned int* node_cnt;
    int64_t nchunks;
    double *pm_off_hi_w, *pm_off_lo_w;  // writable views of pm_off_* (the few-segment merge kernel scans them itself)
    unsigned long long* dbg;            // dev tool: %globaltimer stamps (null: off)
};

// one merge of two solved ranges by a full warp (shared- or global-memory values): the rounds of pav_merge_search_kary (pav_core.h, which the CPU
// tests run with a loop over lanes) with one lane per probe — keep the two in sync
// left half of the merge search: first p in [a, b-1] whose run joins the pooled block (full warp)
template <class V, class PS, class PM>
__device__ int64_t merge_kary_left(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a,
                                   int64_t b, int64_t c, int64_t hint_lo = -1, int64_t hint_hi = -1,
                                   int hint_stride = RBL_HINT_STRIDE) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int64_t lo = a, hi = b - 1, xlo = b, xhi = c;
    bool first = true;
    if (hint_lo >= a && hint_lo < b && hint_hi > b && hint_hi <= c && lo < hi) {
        // warm start (pav_core.h: pav_merge_search_kary): 32 probes around last z-step's block start
        int64_t rr = 0, rs = 0, re = 0;
        const bool pr = pav_probe_left_near(loss, rho, val, ps, pm, a, b, c, pav_hint_pos(hint_lo, lo, hi, lane, hint_stride),
                                            hint_hi, &rr, &rs, &re);
        const unsigned ball = __ballot_sync(FULL, pr);
        const int f = ball ? (__ffs(ball) - 1) : -1;  // first true lane
        const int fs = f >= 0 ? f : 31, fm = f > 0 ? f - 1 : 0;
        const int64_t rs_f = __shfl_sync(FULL, rs, fs), rr_f = __shfl_sync(FULL, rr, fs), re_f = __shfl_sync(FULL, re, fs);
        const int64_t re_m = __shfl_sync(FULL, re, fm), rr_m = __shfl_sync(FULL, rr, fm);
        if (f >= 0) {
            hi = rs_f < hi ? rs_f : hi;
            xhi = rr_f;
            if (f > 0) {
                lo = re_m > lo ? re_m : lo;
                xlo = rr_m;
            }
        } else {
            lo = re_f > lo ? re_f : lo;
            xlo = rr_f;
        }
        if (hi < lo) hi = lo;
        first = false;
    }
    while (lo < hi) {
        const int64_t width = hi - lo;
        int active;
        int64_t pos;
        if (first) {
            active = 0;
            while (active < 32 && ((int64_t)1 << active) <= width) ++active;
            pos = hi - ((int64_t)1 << lane);
        } else {
            active = width <= 32 ? (int)width : 32;
            pos = pav_kary_pos(lo, width, lane);
        }
        int64_t rr = 0, rs = 0, re = 0;
        bool pr = first;  // idle lanes must not look like the searched-for transition
        if (lane < active) pr = pav_probe_left(loss, rho, val, ps, pm, a, b, pos, xlo, xhi, first, &rr, &rs, &re);
        const unsigned ball = __ballot_sync(FULL, pr);
        if (first) {
            const unsigned nb = ~ball;  // first false lane (idle lanes vote true)
            const int f = nb ? (__ffs(nb) - 1) : active;
            const int fs = f < 32 ? f : 31, fm = f > 0 ? f - 1 : 0;
            const int64_t re_f = __shfl_sync(FULL, re, fs), rr_f = __shfl_sync(FULL, rr, fs);
            const int64_t rs_m = __shfl_sync(FULL, rs, fm), rr_m = __shfl_sync(FULL, rr, fm);
            if (f < active) {
                lo = re_f > lo ? re_f : lo;
                xlo = rr_f;
            }
            if (f > 0) {
                hi = rs_m < hi ? rs_m : hi;
                xhi = rr_m;
            }
            first = false;
        } else {
            const int f = ball ? (__ffs(ball) - 1) : -1;  // first true lane
            const int fs = f >= 0 ? f : active - 1, fm = f > 0 ? f - 1 : 0;
            const int64_t rs_f = __shfl_sync(FULL, rs, fs), rr_f = __shfl_sync(FULL, rr, fs), re_f = __shfl_sync(FULL, re, fs);
            const int64_t re_m = __shfl_sync(FULL, re, fm), rr_m = __shfl_sync(FULL, rr, fm);
            if (f >= 0) {
                hi = rs_f < hi ? rs_f : hi;
                xhi = rr_f;
                if (f > 0) {
                    lo = re_m > lo ? re_m : lo;
                    xlo = rr_m;
                }
            } else {
                lo = re_f > lo ? re_f : lo;
                xlo = rr_f;
            }
        }
        if (hi < lo) hi = lo;
    }
    return lo;
}

// right half: first p in [b+1, c) whose run stays out of the pooled block, else c (full warp)
template <class V, class PS, class PM>
__device__ int64_t merge_kary_right(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a,
                                    int64_t b, int64_t c, int64_t hint_lo = -1, int64_t hint_hi = -1,
                                    int hint_stride = RBL_HINT_STRIDE) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int64_t lo = b + 1, hi = c, xlo = a, xhi = b;
    bool first = true;
    if (hint_lo >= a && hint_lo < b && hint_hi > b && hint_hi <= c && lo < hi) {
        // warm start: 32 probes around last z-step's block end
        int64_t ll = 0, rs = 0, re = 0;
        const bool pr = pav_probe_right_near(loss, rho, val, ps, pm, a, b, c, pav_hint_pos(hint_hi, lo, hi - 1, lane, hint_stride),
                                             hint_lo, &ll, &rs, &re);
        const unsigned ball = __ballot_sync(FULL, pr);
        const int f = ball ? (__ffs(ball) - 1) : -1;  // first true lane
        const int fs = f >= 0 ? f : 31, fm = f > 0 ? f - 1 : 0;
        const int64_t rs_f = __shfl_sync(FULL, rs, fs), ll_f = __shfl_sync(FULL, ll, fs), re_f = __shfl_sync(FULL, re, fs);
        const int64_t re_m = __shfl_sync(FULL, re, fm), ll_m = __shfl_sync(FULL, ll, fm);
        if (f >= 0) {
            hi = rs_f < hi ? rs_f : hi;
            xhi = ll_f;
            if (f > 0) {
                lo = re_m > lo ? re_m : lo;
                xlo = ll_m;
            }
        } else {
            lo = re_f > lo ? re_f : lo;
            xlo = ll_f;
        }
        if (hi < lo) hi = lo;
        first = false;
    }
    while (lo < hi) {
        const int64_t width = hi - lo;
        int active;
        int64_t pos;
        if (first) {
            active = 0;
            while (active < 32 && ((int64_t)1 << active) <= width) ++active;
            pos = lo + ((int64_t)1 << lane) - 1;
        } else {
            active = width <= 32 ? (int)width : 32;
            pos = pav_kary_pos(lo, width, lane);
        }
        int64_t ll = 0, rs = 0, re = 0;
        bool pr = false;
        if (lane < active) pr = pav_probe_right(loss, rho, val, ps, pm, b, c, pos, xlo, xhi, first, &ll, &rs, &re);
        const unsigned ball = __ballot_sync(FULL, pr);
        const int f = ball ? (__ffs(ball) - 1) : -1;  // first true lane
        const int fs = f >= 0 ? f : active - 1, fm = f > 0 ? f - 1 : 0;
        const int64_t rs_f = __shfl_sync(FULL, rs, fs), ll_f = __shfl_sync(FULL, ll, fs), re_f = __shfl_sync(FULL, re, fs);
        const int64_t re_m = __shfl_sync(FULL, re, fm), ll_m = __shfl_sync(FULL, ll, fm);
        if (f >= 0) {
            hi = rs_f < hi ? rs_f : hi;
            xhi = ll_f;
            if (f > 0) {
                lo = re_m > lo ? re_m : lo;
                xlo = ll_m;
            }
        } else {
            lo = re_f > lo ? re_f : lo;
            xlo = ll_f;
        }
        first = false;
        if (hi < lo) hi = lo;
    }
    return lo;
}

// one merge of two solved ranges by a full warp: left search, right search, block value
template <class V, class PS, class PM>
__device__ bool merge_kary_warp(int loss, double rho, const V& val, const PS& ps, const PM& pm, int64_t a, int64_t b,
                                int64_t c, int64_t* lo_out, int64_t* hi_out, double* v_out) {
    if (!(val(b - 1) > val(b))) return false;
    const int64_t lo_star = merge_kary_left(loss, rho, val, ps, pm, a, b, c);
    const int64_t hi_star = merge_kary_right(loss, rho, val, ps, pm, a, b, c);
    pav_kary_finish(loss, rho, val, ps, pm, a, c, lo_star, hi_star, lo_out, hi_out, v_out);
    return true;
}

// Chunk stage: each CTA solves its 1024-element chunk in shared memory (element prox + the 10
// in-chunk merge levels) and writes the chunk-stage block values.
__global__ void __launch_bounds__(kPavThreads, 3) pav_chunk_kernel(const TreeParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChunkSmem& S = *reinterpret_cast<ChunkSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int loss = P.loss;
    const double rho = P.scal ? P.scal[0] : P.rho;
    const int64_t n = P.n;
    const int64_t base = (int64_t)blockIdx.x * kChunk;
    const int len = (int)((n - base < kChunk) ? (n - base) : kChunk);

    // ---- stage 1: this chunk, in shared memory
    for (int li = tid; li < len; li += kPavThreads) {
        S.val[li] = rbl_block_prox(loss, P.sigma[base + li], P.m[base + li], rho);
        S.psh[li] = P.ps_loc_hi[base + li];
        S.psl[li] = P.ps_loc_lo[base + li];
        S.pmh[li] = P.pm_loc_hi[base + li];
        S.pml[li] = P.pm_loc_lo[base + li];
    }
    if (tid == 0) {  // closing entries of the chunk-local exclusive prefixes
        const bool full = (len == kChunk);
        S.psh[len] = full ? P.ps_tot_hi[blockIdx.x] : P.ps_loc_hi[base + len];
        S.psl[len] = full ? P.ps_tot_lo[blockIdx.x] : P.ps_loc_lo[base + len];
        S.pmh[len] = full ? P.pm_tot_hi[blockIdx.x] : P.pm_loc_hi[base + len];
        S.pml[len] = full ? P.pm_tot_lo[blockIdx.x] : P.pm_loc_lo[base + len];
    }
    __syncthreads();
    {
        PrefixFlat ps{S.psh, S.psl}, pm{S.pmh, S.pml};
        ValPlain sval{S.val};
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
                        if (pav_merge_search(loss, rho, sval, ps, pm, a, b, c, &lo, &hi, &v))
                            for (int64_t i = lo; i < hi; ++i) S.val[i] = v;
                    }
                }
                __syncthreads();
            } else {
                // few large pairs (<= 16): one warp per pair runs the 32-ary search, the CTA fills
                const int wid = tid >> 5;
                for (int pr = wid; pr < npairs; pr += kPavThreads / 32) {
                    const int a = pr * 2 * w, b = a + w;
                    int64_t lo = 0, hi = 0;
                    double v = 0.0;
                    if (b < len) {
                        const int c = (a + 2 * w < len) ? a + 2 * w : len;
                        if (!merge_kary_warp(loss, rho, sval, ps, pm, a, b, c, &lo, &hi, &v)) lo = hi = 0;
                    }
                    if ((tid & 31) == 0) {
                        S.rec_lo[pr] = lo;
                        S.rec_hi[pr] = hi;
                        S.rec_v[pr] = v;
                    }
                }
                __syncthreads();
                for (int i = tid; i < len; i += kPavThreads) {
                    const int pr = i / (2 * w);
                    if (i >= S.rec_lo[pr] && i < S.rec_hi[pr]) S.val[i] = S.rec_v[pr];
                }
                __syncthreads();
            }
        }
    }
    for (int i = tid; i < len; i += kPavThreads) P.val[base + i] = S.val[i];
}

// Tree stage: one CTA per pair of chunks merges them and then climbs the merge tree: at every node
// the first child to arrive leaves, the second one performs the merge ("last arriver continues" —
// nobody ever waits, so there is no inter-CTA spinning and no co-residency requirement).  Warp 0
// runs the merge search (32-ary, galloping outward from the boundary), the whole CTA fills the
// pooled block.
__global__ void __launch_bounds__(kPavThreads) pav_tree_kernel(const TreeParams P) {
    __shared__ unsigned int s_arrive;
    __shared__ int64_t s_lo, s_hi;
    __shared__ double s_v;
    const int tid = threadIdx.x;
    const int loss = P.loss;
    const double rho = P.scal ? P.scal[0] : P.rho;
    const int64_t n = P.n;
    PrefixChunked gps{P.ps_loc_hi, P.ps_loc_lo, P.ps_off_hi, P.ps_off_lo, kChunkLog2};
    PrefixChunked gpm{P.pm_loc_hi, P.pm_loc_lo, P.pm_off_hi, P.pm_off_lo, kChunkLog2};
    ValCG gval{P.val};
    int64_t parent = blockIdx.x;  // this CTA starts as the only arriver of level-0 node blockIdx.x
    long long cnt_base = 0, level_nodes = (P.nchunks + 1) / 2;
    bool need_arrive = false;
    for (int64_t w = kChunk; w < n; w <<= 1) {
        const int64_t a = parent * 2 * w, b = a + w;
        if (b < n) {  // the node has two children
            if (need_arrive) {
                __threadfence();  // publish this CTA's fills
                __syncthreads();
                if (tid == 0) s_arrive = atomicAdd(&P.node_cnt[cnt_base + parent], 1u);
                __syncthreads();
                if (s_arrive == 0) return;  // first to arrive: the sibling's CTA merges
            }
            if (tid < 32) {
                if (need_arrive) {
                    if (tid == 0) P.node_cnt[cnt_base + parent] = 0;  // self-cleaning for the next call
                    __threadfence();                                   // acquire the sibling's writes
                }
                const int64_t c = (a + 2 * w < n) ? a + 2 * w : n;
                int64_t lo = 0, hi = 0;
                double v = 0.0;
                if (!merge_kary_warp(loss, rho, gval, gps, gpm, a, b, c, &lo, &hi, &v)) lo = hi = 0;
                if (tid == 0) {
                    s_lo = lo;
                    s_hi = hi;
                    s_v = v;
                }
            }
            __syncthreads();
            const int64_t lo = s_lo, hi = s_hi;
            const double v = s_v;
            for (int64_t i = lo + tid; i < hi; i += kPavThreads) P.val[i] = v;
        }
        cnt_base += level_nodes;
        level_nodes = (level_nodes + 1) / 2;
        parent >>= 1;
        need_arrive = true;
    }
}

// ---- few-segment path ------------------------------------------------------------------------------------
constexpr int kMaxSeg = 8;

// value array with a short list of pending pooled blocks laid over it (disjoint, at most kMaxSeg)
struct ValOverlay {
    const double* p;
    const int* nblk;
    const int64_t* lo;
    const int64_t* hi;
    const double* v;
    __device__ __forceinline__ double operator()(int64_t i) const {
        const int nb = *nblk;
        for (int k = 0; k < nb; ++k)
            if (i >= lo[k] && i < hi[k]) return v[k];
        return __ldg(p + i);
    }
};

struct SegBlocks {
    int nblk;
    int pad;
    int64_t lo[kMaxSeg], hi[kMaxSeg];
    double v[kMaxSeg];
    // the pooled block merge j produced at the previous call on this handle (0, 0: none yet) — the warm start of
    // the next call's searches; ranks move little between two ADMM iterations
    int64_t hint_lo[kMaxSeg], hint_hi[kMaxSeg];
    // ... and how far each end moved between the last two calls: along a solve the pooled block shrinks by a slowly
    // decaying number of ranks per iteration (measured at n = 1M: 13 k ranks at iteration 10, 2.6 k at 20, 1 k at
    // 30), so the guess is the last answer extrapolated by 7/8 of its last move
    int64_t move_lo[kMaxSeg], move_hi[kMaxSeg];
    // 1 + |answer - guess| of the last warm-started call (0: not known): while both ends keep landing within 64 ranks
    // of their guesses the first round probes at RBL_HINT_STRIDE_NEAR instead of RBL_HINT_STRIDE
    int64_t err_lo[kMaxSeg], err_hi[kMaxSeg];
};

// positions i (1 <= i < n) where sigma steps up; count may exceed cap (then the list is truncated)
__global__ void sigma_ascents_kernel(const double* __restrict__ sigma, int64_t n, int cap, int* __restrict__ count,
                                     int64_t* __restrict__ pos) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (sigma[i] > sigma[i - 1]) {
            const int k = atomicAdd(count, 1);
            if (k < cap) pos[k] = i;
        }
    }
}

constexpr int kSegMergeThreads = 1024;  // all of them scan the chunk totals; warps 0 and 1 then run the searches

// Shared-memory windows around the two guessed block ends of a merge: the searches, the finish and the block value
// read block values and prefix sums at ~20 dependent positions per round, all within a few hundred ranks of the
// guesses when those are good — from shared memory (~30 cycles) instead of L2 (~600).  Outside the windows the
// accessors fall through to global memory, so a bad guess costs time, never correctness; the cached entries are the
// exact values the global accessors return (the prefix entries are combined with the same dd_add).
constexpr int kWinHalf = 640;  // the warm-started round probes h - 480 .. h + 512
constexpr int kWinLen = 2 * kWinHalf + 1;  // positions per window (prefix sums have an entry one past the end)
struct SegWindows {
    double val[2][kWinLen];
    double psh[2][kWinLen], psl[2][kWinLen], pmh[2][kWinLen], pml[2][kWinLen];
};

template <class V>
struct ValWin {
    V base;
    const double* w[2];
    int64_t b[2], e[2];  // window k caches positions [b[k], e[k])
    __device__ __forceinline__ double operator()(int64_t i) const {
        if (i >= b[0] && i < e[0]) return w[0][i - b[0]];
        if (i >= b[1] && i < e[1]) return w[1][i - b[1]];
        return base(i);
    }
};

struct PrefWin {
    PrefixChunked base;
    const double *h[2], *l[2];
    int64_t b[2], e[2];  // window k caches entries [b[k], e[k]]  (inclusive end)
    __device__ __forceinline__ dd_t get(int64_t i) const {
        dd_t r;
        if (i >= b[0] && i <= e[0]) {
            r.hi = h[0][i - b[0]];
            r.lo = l[0][i - b[0]];
            return r;
        }
        if (i >= b[1] && i <= e[1]) {
            r.hi = h[1][i - b[1]];
            r.lo = l[1][i - b[1]];
            return r;
        }
        return base.get(i);
    }
};

// two warps: merge the solved prefix [0, bounds[j]) with the run [bounds[j], bounds[j+1]) for j = 1..nseg-1;
// warp 0 searches the left end of the pooled block while warp 1 searches the right end (independent
// dependent-load chains, ~50 us each at n = 1M)
__global__ void __launch_bounds__(kSegMergeThreads) pav_seg_merge_kernel(const TreeParams P,
                                                                         const int64_t* __restrict__ bounds, int nseg,
                                                                         SegBlocks* __restrict__ out, int use_hints) {
    rbl_pdl_wait();
    extern __shared__ __align__(16) unsigned char seg_smem[];
    SegWindows& W = *reinterpret_cast<SegWindows*>(seg_smem);
    __shared__ int s_nblk;
    __shared__ int64_t s_lo[kMaxSeg], s_hi[kMaxSeg];
    __shared__ double s_v[kMaxSeg];
    __shared__ int64_t s_end[2], s_snap[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double rho = P.scal ? P.scal[0] : P.rho;
    int dbg_n = 0;
    auto stamp = [&]() {
        if (P.dbg && tid == 0 && dbg_n < 60) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
            P.dbg[1 + dbg_n++] = tt;
        }
    };
    stamp();
    if (tid == 0) s_nblk = 0;
    {   // exclusive scan of the chunk totals of the margins (what chunk_offsets_kernel does for the tree route):
        // the whole CTA takes part (one or a few chunks per thread — 64 threads walking 16 chunks each was a
        // quarter of this kernel's time), double-double throughout, fixed order
        __shared__ double s_wh[32], s_wl[32];
        const int nt = blockDim.x, nw = nt >> 5;
        const int64_t nch = P.nchunks;
        const int64_t per = (nch + nt - 1) / nt;
        const int64_t c0 = (int64_t)tid * per;
        dd_t run = dd_make(0.0);
        for (int64_t c = c0; c < c0 + per && c < nch; ++c) {
            dd_t t;
            t.hi = P.pm_tot_hi[c];
            t.lo = P.pm_tot_lo[c];
            run = dd_add(run, t);
        }
        dd_t x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            dd_t y = shfl_up_dd(x, o);
            if (lane >= o) x = dd_add(y, x);
        }
        if (lane == 31) {
            s_wh[warp] = x.hi;
            s_wl[warp] = x.lo;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the warp totals
            dd_t t = dd_make(0.0);
            if (lane < nw) {
                t.hi = s_wh[lane];
                t.lo = s_wl[lane];
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                dd_t y = shfl_up_dd(t, o);
                if (lane >= o) t = dd_add(y, t);
            }
            dd_t e = shfl_up_dd(t, 1);
            if (lane == 0) e = dd_make(0.0);
            s_wh[lane] = e.hi;
            s_wl[lane] = e.lo;
        }
        __syncthreads();
        dd_t incl_prev = shfl_up_dd(x, 1);
        if (lane == 0) incl_prev = dd_make(0.0);
        dd_t wbase;
        wbase.hi = s_wh[warp];
        wbase.lo = s_wl[warp];
        dd_t ex = dd_add(wbase, incl_prev);
        for (int64_t c = c0; c < c0 + per && c < nch; ++c) {
            P.pm_off_hi_w[c] = ex.hi;
            P.pm_off_lo_w[c] = ex.lo;
            dd_t t;
            t.hi = P.pm_tot_hi[c];
            t.lo = P.pm_tot_lo[c];
            ex = dd_add(ex, t);
        }
        if (tid == nt - 1) {  // grand total closes the offsets (the last thread owns the last slice, possibly empty)
            P.pm_off_hi_w[nch] = ex.hi;
            P.pm_off_lo_w[nch] = ex.lo;
        }
    }
    __syncthreads();  // offsets (global memory, this CTA's own writes) and s_nblk visible
    stamp();
    PrefixChunked gps{P.ps_loc_hi, P.ps_loc_lo, P.ps_off_hi, P.ps_off_lo, kChunkLog2};
    PrefixChunked gpm{P.pm_loc_hi, P.pm_loc_lo, P.pm_off_hi, P.pm_off_lo, kChunkLog2};
    ValOverlay gval{P.val, &s_nblk, s_lo, s_hi, s_v};
    for (int j = 1; j < nseg; ++j) {
        const int64_t b = bounds[j], c = bounds[j + 1];
        const bool violated = gval(b - 1) > gval(b);  // block-uniform
        const int64_t p_lo = use_hints ? out->hint_lo[j] : -1, p_hi = use_hints ? out->hint_hi[j] : -1;
        const bool have_prev = p_lo >= 0 && p_lo < b && p_hi > b && p_hi <= c;
        int64_t h_lo = -1, h_hi = -1;
        if (have_prev) {  // last answer + 7/8 of its last move, kept inside the valid ranges
            h_lo = p_lo + (out->move_lo[j] * 7) / 8;
            h_hi = p_hi + (out->move_hi[j] * 7) / 8;
            h_lo = h_lo < 0 ? 0 : (h_lo > b - 1 ? b - 1 : h_lo);
            h_hi = h_hi < b + 1 ? b + 1 : (h_hi > c ? c : h_hi);
        }
        const bool hinted = have_prev;
        int stride = RBL_HINT_STRIDE;
        if (have_prev) {
            const int64_t el = out->err_lo[j], eh = out->err_hi[j];
            if (el > 0 && el <= 65 && eh > 0 && eh <= 65) stride = RBL_HINT_STRIDE_NEAR;
        }
        // windows around the guessed block ends (empty without a usable guess): filled by the whole CTA
        ValWin<ValOverlay> val{gval, {W.val[0], W.val[1]}, {0, 0}, {0, 0}};
        PrefWin ps{gps, {W.psh[0], W.psh[1]}, {W.psl[0], W.psl[1]}, {0, 0}, {-1, -1}};
        PrefWin pm{gpm, {W.pmh[0], W.pmh[1]}, {W.pml[0], W.pml[1]}, {0, 0}, {-1, -1}};
        if (violated && hinted) {
            for (int k = 0; k < 2; ++k) {
                const int64_t ctr = k == 0 ? h_lo : h_hi;
                int64_t w0 = ctr - kWinHalf, w1 = ctr + kWinHalf;  // entries [w0, w1] of the prefixes, [w0, w1) of val
                if (w0 < 0) w0 = 0;
                if (w1 > P.n) w1 = P.n;
                if (k == 1 && w0 <= val.e[0] && val.e[0] > val.b[0]) w0 = val.e[0] < w1 ? val.e[0] : w1;  // disjoint
                val.b[k] = w0;
                val.e[k] = w1;
                ps.b[k] = pm.b[k] = w0;
                ps.e[k] = pm.e[k] = w1;
                for (int64_t i = w0 + tid; i <= w1; i += kSegMergeThreads) {
                    // all nine loads of a position are issued before any is used (PrefixChunked::get would make
                    // the chunk-local entries wait for the offsets); same dd_add, same values
                    const int64_t ch = i >> kChunkLog2;
                    const bool inner = (i & (((int64_t)1 << kChunkLog2) - 1)) != 0;
                    const double v0 = i < w1 ? __ldg(P.val + i) : 0.0;
                    dd_t so, sl, mo, ml;
                    so.hi = P.ps_off_hi[ch];
                    so.lo = P.ps_off_lo[ch];
                    mo.hi = P.pm_off_hi[ch];
                    mo.lo = P.pm_off_lo[ch];
                    sl.hi = P.ps_loc_hi[i];
                    sl.lo = P.ps_loc_lo[i];
                    ml.hi = P.pm_loc_hi[i];
                    ml.lo = P.pm_loc_lo[i];
                    if (i < w1) {
                        double vv = v0;
                        const int nb = s_nblk;
                        for (int q = 0; q < nb; ++q)
                            if (i >= s_lo[q] && i < s_hi[q]) vv = s_v[q];
                        W.val[k][i - w0] = vv;
                    }
                    const dd_t a = inner ? dd_add(so, sl) : so, m2 = inner ? dd_add(mo, ml) : mo;
                    W.psh[k][i - w0] = a.hi;
                    W.psl[k][i - w0] = a.lo;
                    W.pmh[k][i - w0] = m2.hi;
                    W.pml[k][i - w0] = m2.lo;
                }
            }
        }
        __syncthreads();
        stamp();
        if (violated && warp < 2) {
            const int64_t e = warp == 0 ? merge_kary_left(P.loss, rho, val, ps, pm, (int64_t)0, b, c, h_lo, h_hi, stride)
                                        : merge_kary_right(P.loss, rho, val, ps, pm, (int64_t)0, b, c, h_lo, h_hi, stride);
            if (lane == 0) {
                // each side snaps its own answer to a whole run of equal values (what pav_kary_finish does first),
                // the two sides side by side instead of one after the other on thread 0
                s_end[warp] = e;
                s_snap[warp] = warp == 0 ? pav_run_start(val, e, (int64_t)0, val(e))
                                         : pav_run_end(val, e - 1, c, val(e - 1));
            }
        }
        __syncthreads();
        stamp();
        if (violated && tid == 0) {
            const int64_t lo = s_snap[0], hi = s_snap[1];
            double v = pav_block_value(P.loss, rho, ps, pm, lo, hi);  // the rest of pav_kary_finish
            stamp();
            if (lo > 0) {
                const double vl = val(lo - 1);
                if (v < vl) v = vl;
            }
            if (hi < c) {
                const double vr = val(hi);
                if (v > vr) v = vr;
            }
            if (P.dbg && j < 4) {  // how far the snaps moved the searches' answers
                P.dbg[56 + 2 * j] = (unsigned long long)(s_end[0] - lo);
                P.dbg[57 + 2 * j] = (unsigned long long)(hi - s_end[1]);
            }
            if (P.dbg && j < 4) {  // dev tool: guesses and answers of this merge
                P.dbg[40 + 4 * j] = (unsigned long long)h_lo;
                P.dbg[41 + 4 * j] = (unsigned long long)h_hi;
                P.dbg[42 + 4 * j] = (unsigned long long)s_end[0];
                P.dbg[43 + 4 * j] = (unsigned long long)s_end[1];
            }
            out->move_lo[j] = have_prev ? s_end[0] - p_lo : 0;
            out->move_hi[j] = have_prev ? s_end[1] - p_hi : 0;
            out->err_lo[j] = have_prev ? 1 + (s_end[0] > h_lo ? s_end[0] - h_lo : h_lo - s_end[0]) : 0;
            out->err_hi[j] = have_prev ? 1 + (s_end[1] > h_hi ? s_end[1] - h_hi : h_hi - s_end[1]) : 0;
            out->hint_lo[j] = s_end[0];   // (the searches' own answers: the finish only snaps them to whole runs)
            out->hint_hi[j] = s_end[1];
            // blocks are swallowed whole (a probe decides for the whole run of equal values around it)
            int k2 = 0;
            const int nb = s_nblk;
            for (int k = 0; k < nb; ++k) {
                if (s_lo[k] >= lo && s_hi[k] <= hi) continue;
                s_lo[k2] = s_lo[k];
                s_hi[k2] = s_hi[k];
                s_v[k2] = s_v[k];
                ++k2;
            }
            s_lo[k2] = lo;
            s_hi[k2] = hi;
            s_v[k2] = v;
            s_nblk = k2 + 1;
        }
        __syncthreads();
        stamp();
    }
    if (P.dbg && tid == 0) P.dbg[0] = (unsigned long long)dbg_n;
    if (tid == 0) {
        out->nblk = s_nblk;
        for (int k = 0; k < s_nblk; ++k) {
            out->lo[k] = s_lo[k];
            out->hi[k] = s_hi[k];
            out->v[k] = s_v[k];
        }
    }
}

__global__ void pav_seg_fill_kernel(const SegBlocks* __restrict__ blk, double* __restrict__ val) {
    rbl_pdl_wait();
    const int nb = blk->nblk;
    for (int k = 0; k < nb; ++k) {
        const int64_t lo = blk->lo[k], hi = blk->hi[k];
        const double v = blk->v[k];
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
             i += (int64_t)gridDim.x * blockDim.x)
            val[i] = v;
    }
}

// element-wise prox without pooling (PAV level 0 as a standalone op; individual_solver.py:112-130)
__global__ void prox_elementwise_kernel(int loss, double rho, const double* __restrict__ sigma,
                                        const double* __restrict__ m, int64_t n, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = rbl_block_prox(loss, sigma[i], m[i], rho);
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
    RBL_CUDA(rbl_launch_pdl(chunk_prefix_kernel, dim3((unsigned)nch), dim3(kPavThreads), 0, s, x, n, loc_hi, loc_lo, tot_hi, tot_lo, nullptr, 0, 0.0,
                                                             nullptr, nullptr));
    RBL_LAUNCH_CHECK();
    chunk_offsets_kernel<<<1, kPavThreads, 0, s>>>(tot_hi, tot_lo, nch, off_hi, off_lo);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_pav_max_seg() { return kMaxSeg; }
size_t rbl_pav_segblocks_bytes() { return sizeof(SegBlocks); }

// runs of non-increasing sigma: writes the ascending run boundaries [0, ..., n] to c->seg_bounds (device) and
// sets c->nseg (0: more than kMaxSeg runs, use the merge tree).  Synchronises `s` (set-up time only).
int rbl_k_segments(rbl_ctx* c, cudaStream_t s) {
    const int cap = kMaxSeg;
    int count = 0;
    int64_t pos[kMaxSeg + 2];
    RBL_CUDA(cudaMemsetAsync(c->seg_count, 0, sizeof(int), s));
    sigma_ascents_kernel<<<c->vec_grid, 256, 0, s>>>(c->sigma, c->n_global, cap, c->seg_count, c->seg_bounds + 1);
    RBL_LAUNCH_CHECK();
    RBL_CUDA(cudaMemcpyAsync(&count, c->seg_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    RBL_CUDA(cudaStreamSynchronize(s));
    if (count + 1 > kMaxSeg) {
        c->nseg = 0;
        return RBL_OK;
    }
    RBL_CUDA(cudaMemcpy(pos + 1, c->seg_bounds + 1, (size_t)count * sizeof(int64_t), cudaMemcpyDeviceToHost));
    for (int i = 1; i <= count; ++i)  // insertion sort (atomics filled the list in arbitrary order)
        for (int j = i; j > 1 && pos[j] < pos[j - 1]; --j) {
            const int64_t t = pos[j];
            pos[j] = pos[j - 1];
            pos[j - 1] = t;
        }
    pos[0] = 0;
    pos[count + 1] = c->n_global;
    RBL_CUDA(cudaMemcpy(c->seg_bounds, pos, (size_t)(count + 2) * sizeof(int64_t), cudaMemcpyHostToDevice));
    c->nseg = count + 1;
    return RBL_OK;
}

// z_sorted = isotonic prox of the sorted margins: 2 prefix kernels + 1 tree kernel
int rbl_k_pav(rbl_ctx* c, int loss, const double* m_sorted, double rho, double* z_sorted, cudaStream_t s) {
    const int64_t n = c->n_global;
    const int64_t nch = c->nchunks;
    RBL_PER_DEVICE(bool, attr_set, c);
    if (!attr_set) {
        RBL_CUDA(cudaFuncSetAttribute(pav_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(ChunkSmem)));
        attr_set = true;
    }
    const bool few = c->nseg > 0 && !c->force_tree;
    RBL_CUDA(rbl_launch_pdl(chunk_prefix_kernel, dim3((unsigned)nch), dim3(kPavThreads), 0, s, m_sorted, n, c->pm_loc_hi, c->pm_loc_lo, c->ch_tot_hi,
                                                             c->ch_tot_lo, c->sigma, loss, rho,
                                                             few ? z_sorted : nullptr, c->scal));
    RBL_LAUNCH_CHECK();
    if (few && c->nseg == 1) return RBL_OK;  // sigma never steps up (ERM): the element prox is the answer
    if (!few) {
        chunk_offsets_kernel<<<1, kPavThreads, 0, s>>>(c->ch_tot_hi, c->ch_tot_lo, nch, c->pm_off_hi, c->pm_off_lo);
        RBL_LAUNCH_CHECK();
    }
    TreeParams P;
    P.loss = loss;
    P.rho = rho;
    P.scal = c->scal;
    P.sigma = c->sigma;
    P.m = m_sorted;
    P.n = n;
    P.ps_loc_hi = c->ps_loc_hi; P.ps_loc_lo = c->ps_loc_lo;
    P.ps_tot_hi = c->ps_tot_hi; P.ps_tot_lo = c->ps_tot_lo;
    P.ps_off_hi = c->ps_off_hi; P.ps_off_lo = c->ps_off_lo;
    P.pm_loc_hi = c->pm_loc_hi; P.pm_loc_lo = c->pm_loc_lo;
    P.pm_tot_hi = c->ch_tot_hi; P.pm_tot_lo = c->ch_tot_lo;
    P.pm_off_hi = c->pm_off_hi; P.pm_off_lo = c->pm_off_lo;
    P.val = z_sorted;
    P.node_cnt = c->node_cnt;
    P.nchunks = nch;
    P.pm_off_hi_w = c->pm_off_hi;
    P.pm_off_lo_w = c->pm_off_lo;
    P.dbg = c->sort_dbg;
    if (few) {
        SegBlocks* blk = reinterpret_cast<SegBlocks*>(c->seg_blocks);
        RBL_PER_DEVICE(bool, seg_attr, c);
        if (!seg_attr) {
            RBL_CUDA(cudaFuncSetAttribute(pav_seg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(SegWindows)));
            seg_attr = true;
        }
        RBL_CUDA(rbl_launch_pdl(pav_seg_merge_kernel, dim3(1), dim3(kSegMergeThreads), sizeof(SegWindows), s, P,
                                c->seg_bounds, c->nseg, blk, c->pav_no_hints ? 0 : 1));
        RBL_LAUNCH_CHECK();
        RBL_CUDA(rbl_launch_pdl(pav_seg_fill_kernel, dim3(c->vec_grid), dim3(256), 0, s, blk, z_sorted));
        RBL_LAUNCH_CHECK();
        return RBL_OK;
    }
    pav_chunk_kernel<<<(unsigned)nch, kPavThreads, sizeof(ChunkSmem), s>>>(P);
    RBL_LAUNCH_CHECK();
    if (nch > 1) {
        pav_tree_kernel<<<(unsigned)((nch + 1) / 2), kPavThreads, 0, s>>>(P);
        RBL_LAUNCH_CHECK();
    }
    return RBL_OK;
}
