// pass_kernels.cu — one streaming pass over the n x d design matrix D = -y (.) X (row-major fp64).
//
// Replaces the dense matvecs of the reference's w-step and margins:
//   D @ w                      src/optim/algorithms.py:89,132,135
//   y - X b_p ; X^T r ; y - X b src/util/fast_lasso.py:41-56   (three passes per FISTA iteration)
//   D @ w - b ; D^T(.)         src/util/w_LBFGS.py:31-45
// with ONE kernel that, per row tile, forms the row dots d_i = D_i . x, the residual
// r_i = b_i - d_i, its squared norm, and the column accumulation g += D_i^T r_i — so D is read
// from HBM exactly once per FISTA trial / L-BFGS evaluation (fp64 matvec: 0.25 flop/byte, HBM bound;
// no tensor cores).
//
// B200 mapping:
//   * persistent grid, one CTA per SM, 256 threads; tiles of R consecutive rows are a single
//     contiguous R*ld*8-byte span of HBM, fetched with ONE cp.async.bulk (TMA bulk copy, UBLKCP)
//     into a 2-4 stage shared-memory ring tracked by mbarriers (expect_tx / complete_tx);
//     128-192 KB in flight per SM, far above latency x bandwidth (~35 KB/SM at 6.5 TB/s).
//   * phase A (row dots): a warp (or WPR warps for d > 1024) owns a row, lanes read 128-bit
//     double2 from shared memory, x lives in registers, warp-shuffle reduction.
//   * phase B (column accumulation): each thread owns fixed double2 columns, accumulators in
//     registers across ALL tiles of the CTA; conflict-free 128-bit shared-memory reads.
//   * per-CTA partials are written once at the end and summed in a fixed order by
//     rbl_reduce_partials (deterministic, and identical on every rank after the all-reduce).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// storage type of the design matrix: fp64 (default) or fp32 (optional mode, rbl_set_storage: halves the HBM bytes of
// every pass; products and sums stay fp64)
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };
__device__ __forceinline__ double2 to_d2(double2 v) { return v; }
__device__ __forceinline__ double2 to_d2(float2 v) { return make_double2((double)v.x, (double)v.y); }

struct PassParams {
    const void* D;
    int64_t ld;
    int64_t n;
    int d;
    const double* x;
    const double* b;
    double* out;
    double* gpart;
    double* sspart;
    const FistaState* st;
    double* r0;
    double* r1;
    int R;
    int stages;
    int mode;
    int evict_first;
    uint32_t stage_stride;  // bytes, multiple of 128
    double* lam;            // RBL_PASS_DUAL: multiplier updated in place, b = z
    double rho;
    const double* scal;     // device scalar block overriding rho when bound (may be null)
    const int* gate_nnz;    // RBL_PASS_DUAL: run only if *gate_nnz > gate_cap (w too dense for the sparse path)
    int gate_cap;
};

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxRows = 64;

// T : storage type of D;  XJ : double2 of x held per lane;  WPR : warps cooperating on one row;  CJ : column pairs
// per thread
template <typename T, int XJ, int WPR, int CJ>
__global__ void __launch_bounds__(kThreads, 1) rbl_pass_kernel(const PassParams p) {
    using V2 = typename Vec2<T>::type;
    rbl_pdl_wait();
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int mode = p.mode;
    double* out = p.out;
    if (mode == RBL_PASS_FISTA) {
        if (p.st->done) return;  // converged earlier in this batch of enqueued steps
        out = p.st->cur ? p.r1 : p.r0;
    }
    // gated modes: the sparse-w kernel (DUAL) or the active-row gather (FUSED) took this step
    if ((mode == RBL_PASS_DUAL || mode == RBL_PASS_FUSED) && p.gate_nnz && *p.gate_nnz <= p.gate_cap) return;
    const bool fused = (mode == RBL_PASS_FUSED || mode == RBL_PASS_FISTA);
    const bool dual = (mode == RBL_PASS_DUAL);
    const double rho = p.scal ? p.scal[0] : p.rho;
    const int R = p.R, S = p.stages;
    const int64_t ld = p.ld;
    const int ld2 = (int)(ld >> 1);

    unsigned char* stage_base = smem;
    double* rt = reinterpret_cast<double*>(smem + (size_t)S * p.stage_stride);  // [kMaxRows]
    double* part = rt + kMaxRows;                                                // [kMaxRows * 8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + kMaxRows * 8);           // [S]

    const int64_t ntiles = (p.n + R - 1) / R;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t nmine = ntiles > first ? (ntiles - first + stride - 1) / stride : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint64_t policy = 0;
    if (p.evict_first) policy = l2_policy_evict_first();
    auto issue = [&](int64_t k) {
        const int s = (int)(k % S);
        const int64_t row0 = (first + k * stride) * R;
        const int rows = (int)((p.n - row0 < R) ? (p.n - row0) : R);
        const uint32_t bytes = (uint32_t)((size_t)rows * ld * sizeof(T));
        mbar_expect_tx(&bars[s], bytes);
        void* dst = stage_base + (size_t)s * p.stage_stride;
        const void* src = reinterpret_cast<const T*>(p.D) + row0 * ld;
        if (p.evict_first) tma_bulk_g2s_hint(dst, src, bytes, &bars[s], policy);
        else tma_bulk_g2s(dst, src, bytes, &bars[s]);
    };
    if (tid == 0) {
        const int64_t pre = nmine < S ? nmine : S;
        for (int64_t k = 0; k < pre; ++k) issue(k);
    }

    // x slice of this lane in registers; zero beyond d so padded columns contribute nothing
    const int h = warp % WPR, grp = warp / WPR;
    constexpr int G = kWarps / WPR;
    double2 xr[XJ];
#pragma unroll
    for (int j = 0; j < XJ; ++j) {
        const int c = 2 * (lane + 32 * (j * WPR + h));
        xr[j].x = (c < p.d) ? p.x[c] : 0.0;
        xr[j].y = (c + 1 < p.d) ? p.x[c + 1] : 0.0;
    }
    double2 acc[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[j] = make_double2(0.0, 0.0);
    double ss = 0.0;

    for (int64_t k = 0; k < nmine; ++k) {
        const int s = (int)(k % S);
        const uint32_t parity = (uint32_t)((k / S) & 1);
        const int64_t row0 = (first + k * stride) * R;
        const int rows = (int)((p.n - row0 < R) ? (p.n - row0) : R);
        // issue the b load now; it is consumed after phase A
        double bv = 0.0;
        double lv = 0.0;
        if ((fused || dual) && tid < rows) bv = p.b[row0 + tid];
        if (dual && tid < rows) lv = p.lam[row0 + tid];
        while (!mbar_try_wait(&bars[s], parity)) {
        }
        const V2* T2 = reinterpret_cast<const V2*>(stage_base + (size_t)s * p.stage_stride);

        // ---- phase A: row dots
        for (int i = grp; i < rows; i += G) {
            const V2* row = T2 + (size_t)i * ld2;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < XJ; ++j) {
                const int c2 = lane + 32 * (j * WPR + h);
                if (c2 < ld2) {
                    const double2 v = to_d2(row[c2]);
                    a0 = fma(v.x, xr[j].x, a0);
                    a1 = fma(v.y, xr[j].y, a1);
                }
            }
            double a = a0 + a1;
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            if (lane == 0) part[i * WPR + h] = a;
        }
        __syncthreads();
        if (tid < rows) {
            double dot = part[tid * WPR];
#pragma unroll
            for (int hh = 1; hh < WPR; ++hh) dot += part[tid * WPR + hh];
            if (fused) {
                const double r = bv - dot;
                out[row0 + tid] = r;
                rt[tid] = r;
                ss = fma(r, r, ss);
            } else if (dual) {
                // lambda += rho (z - D w), partial ||z - D w||^2 (algorithms.py:132,135) in the matvec epilogue
                const double res = bv - dot;
                out[row0 + tid] = dot;
                // product and sum rounded separately, like numpy's lam + rho * (z - Dw) (:132): an FMA here leaves
                // 1-ulp residues where the reference cancels to exactly 0 (rows with z = m while w = 0)
                p.lam[row0 + tid] = __dadd_rn(lv, __dmul_rn(rho, res));
                ss = fma(res, res, ss);
            } else {
                out[row0 + tid] = dot;
            }
        }
        if (fused) {
            __syncthreads();
            // ---- phase B: g += D_tile^T r_tile, thread-owned columns
            for (int i = 0; i < rows; ++i) {
                const double ri = rt[i];
                const V2* row = T2 + (size_t)i * ld2;
#pragma unroll
                for (int j = 0; j < CJ; ++j) {
                    const int c2 = tid + kThreads * j;
                    if (c2 < ld2) {
                        const double2 v = to_d2(row[c2]);
                        acc[j].x = fma(v.x, ri, acc[j].x);
                        acc[j].y = fma(v.y, ri, acc[j].y);
                    }
                }
            }
        }
        __syncthreads();  // every thread is done reading stage s
        if (tid == 0 && k + S < nmine) issue(k + S);
    }

    if (fused) {
        double2* gp = reinterpret_cast<double2*>(p.gpart + (size_t)blockIdx.x * ld);
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c2 = tid + kThreads * j;
            if (c2 < ld2) gp[c2] = acc[j];
        }
    }
    if (fused || dual) {
        // block-reduce ss (only threads < R hold non-zero values); fixed order
        ss += __shfl_xor_sync(0xffffffffu, ss, 16);
        ss += __shfl_xor_sync(0xffffffffu, ss, 8);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        __syncthreads();
        if (lane == 0) part[warp] = ss;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kWarps; ++w) t += part[w];
            p.sspart[blockIdx.x] = t;
        }
    }
}

// ---- gather pass: g = sum_k delta_k D[row_k] over the ACTIVE rows only (vec_kernels.cu: scatter_active) -------
// Same persistent grid and shared-memory ring as rbl_pass_kernel, but a tile is R rows picked from the list:
// lanes 0..R-1 of warp 0 each issue one cp.async.bulk (a whole 8 d-byte row) against the stage's mbarrier, so
// HBM still sees multi-KB contiguous bursts.  Only the column accumulation runs (r = delta is known), and the
// traffic is (#active rows / n) of a full pass: 0.2 for superquantile q = 0.8 once the pooled block is small.
struct GatherParams {
    const void* D;
    int64_t ld;
    const int32_t* rows;
    const double* delta;
    const int* count;
    int cap;
    double* gpart;
    double* sspart;
    int R;
    int stages;
    uint32_t stage_stride;
};

template <typename T, int CJ>
__global__ void __launch_bounds__(kThreads, 1) rbl_gather_kernel(const GatherParams p) {
    using V2 = typename Vec2<T>::type;
    rbl_pdl_wait();
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int count = *p.count;
    if (count > p.cap) return;  // too dense: the streaming pass takes it
    const int R = p.R, S = p.stages;
    const int64_t ld = p.ld;
    const int ld2 = (int)(ld >> 1);
    unsigned char* stage_base = smem;
    double* dl = reinterpret_cast<double*>(smem + (size_t)S * p.stage_stride);  // [4][kMaxRows] (S <= 4)
    uint64_t* bars = reinterpret_cast<uint64_t*>(dl + 4 * kMaxRows);             // [S]; inside the pass kernel's
                                                                                 // misc area (9 kMaxRows doubles)

    const int64_t ntiles = ((int64_t)count + R - 1) / R;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t nmine = ntiles > first ? (ntiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint64_t policy = l2_policy_evict_first();
    const uint32_t row_bytes = (uint32_t)(ld * sizeof(T));
    double ss = 0.0;
    // warp 0 issues tile k: lane i < rows copies list entry k R + i
    auto issue = [&](int64_t k) {
        const int s = (int)(k % S);
        const int64_t e0 = (first + k * stride) * R;
        const int rows = (int)((count - e0 < R) ? (count - e0) : R);
        int32_t row = 0;
        double dv = 0.0;
        if (lane < rows) {
            row = p.rows[e0 + lane];
            dv = p.delta[e0 + lane];
        }
        if (lane == 0) mbar_expect_tx(&bars[s], (uint32_t)rows * row_bytes);
        __syncwarp();
        if (lane < rows) {
            tma_bulk_g2s_hint(stage_base + (size_t)s * p.stage_stride + (size_t)lane * row_bytes,
                              reinterpret_cast<const T*>(p.D) + (int64_t)row * ld, row_bytes, &bars[s], policy);
            dl[s * kMaxRows + lane] = dv;
            ss = fma(dv, dv, ss);
        }
    };
    if (warp == 0) {
        const int64_t pre = nmine < S ? nmine : S;
        for (int64_t k = 0; k < pre; ++k) issue(k);
    }
    __syncthreads();  // dl of the first stages visible

    double2 acc[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[j] = make_double2(0.0, 0.0);
    for (int64_t k = 0; k < nmine; ++k) {
        const int s = (int)(k % S);
        const uint32_t parity = (uint32_t)((k / S) & 1);
        const int64_t e0 = (first + k * stride) * R;
        const int rows = (int)((count - e0 < R) ? (count - e0) : R);
        while (!mbar_try_wait(&bars[s], parity)) {
        }
        const V2* T2 = reinterpret_cast<const V2*>(stage_base + (size_t)s * p.stage_stride);
        for (int i = 0; i < rows; ++i) {
            const double ri = dl[s * kMaxRows + i];
            const V2* row = T2 + (size_t)i * ld2;
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
                const int c2 = tid + kThreads * j;
                if (c2 < ld2) {
                    const double2 v = to_d2(row[c2]);
                    acc[j].x = fma(v.x, ri, acc[j].x);
                    acc[j].y = fma(v.y, ri, acc[j].y);
                }
            }
        }
        __syncthreads();  // every thread is done with stage s (rows and dl)
        if (warp == 0 && k + S < nmine) issue(k + S);
        __syncthreads();  // dl of the re-issued stage visible before anyone reaches it
    }
    double2* gp = reinterpret_cast<double2*>(p.gpart + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int c2 = tid + kThreads * j;
        if (c2 < ld2) gp[c2] = acc[j];
    }
    // ||delta||^2 of my tiles: lanes of warp 0 hold the terms, fixed-order reduction
    if (warp == 0) {
        ss += __shfl_xor_sync(0xffffffffu, ss, 16);
        ss += __shfl_xor_sync(0xffffffffu, ss, 8);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        if (lane == 0) p.sspart[blockIdx.x] = ss;
    }
}

template <typename T, int CJ>
int launch_gather_tt(rbl_ctx* c, const GatherParams& p, cudaStream_t s) {
    RBL_PER_DEVICE(size_t, attr_smem, c);
    if (c->pass_smem > attr_smem) {
        RBL_CUDA(cudaFuncSetAttribute(rbl_gather_kernel<T, CJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)c->pass_smem));
        attr_smem = c->pass_smem;
    }
    RBL_CUDA(rbl_launch_pdl(rbl_gather_kernel<T, CJ>, dim3(c->pass_grid), dim3(kThreads), c->pass_smem, s, p));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

template <int CJ>
int launch_gather_t(rbl_ctx* c, const GatherParams& p, cudaStream_t s) {
    return c->esz == 4 ? launch_gather_tt<float, CJ>(c, p, s) : launch_gather_tt<double, CJ>(c, p, s);
}

template <typename T, int XJ, int WPR, int CJ>
int launch_tt(rbl_ctx* c, const PassParams& p, cudaStream_t s) {
    RBL_PER_DEVICE(size_t, attr_smem, c);  // per instantiation and per device
    if (c->pass_smem > attr_smem) {
        RBL_CUDA(cudaFuncSetAttribute(rbl_pass_kernel<T, XJ, WPR, CJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)c->pass_smem));
        attr_smem = c->pass_smem;
    }
    RBL_CUDA(rbl_launch_pdl(rbl_pass_kernel<T, XJ, WPR, CJ>, dim3(c->pass_grid), dim3(kThreads), c->pass_smem, s, p));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

template <int XJ, int WPR, int CJ>
int launch_t(rbl_ctx* c, const PassParams& p, cudaStream_t s) {
    return c->esz == 4 ? launch_tt<float, XJ, WPR, CJ>(c, p, s) : launch_tt<double, XJ, WPR, CJ>(c, p, s);
}

int pick_wpr(int64_t ld2) { return ld2 <= 512 ? 1 : ld2 <= 1024 ? 2 : ld2 <= 2048 ? 4 : 8; }

}  // namespace

int rbl_pass_configure(rbl_ctx* c) {
    const int64_t ld2 = c->ld / 2;
    if (c->esz != 4) c->esz = 8;
    if ((c->ld * c->esz) % 16 != 0) {
        rbl_set_error("rows must be a multiple of 16 bytes for the TMA bulk copies: leading dimension %lld must be a "
                      "multiple of %d for %d-byte elements", (long long)c->ld, 16 / c->esz, c->esz);
        return RBL_ERR_ARG;
    }
    if (ld2 > 4096) {
        rbl_set_error("d = %d > 8192 is not supported by the row-tile pass kernel yet", c->d);
        return RBL_ERR_UNSUPPORTED;
    }
    const int wpr = pick_wpr(ld2);
    const int G = kWarps / wpr;
    const size_t row_bytes = (size_t)c->ld * c->esz;
    int R = (int)(65536 / row_bytes);
    R = (R / G) * G;
    if (R < G) R = G;
    if (R > kMaxRows) R = kMaxRows;
    const size_t stage = ((size_t)R * row_bytes + 127) & ~(size_t)127;
    const size_t misc = (kMaxRows + kMaxRows * 8) * sizeof(double) + 8 * sizeof(uint64_t);
    int dev_max = 0;
    RBL_CUDA(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    int stages = (int)(((size_t)dev_max - misc) / stage);
    if (stages > 4) stages = 4;
    if (stages < 1) {
        rbl_set_error("row tile of %d rows x %lld doubles does not fit in shared memory", R, (long long)c->ld);
        return RBL_ERR_UNSUPPORTED;
    }
    c->pass_rows = R;
    c->pass_stages = stages;
    c->pass_smem = (size_t)stages * stage + misc;
    c->pass_grid = c->num_sms;
    return RBL_OK;
}

int rbl_launch_pass(rbl_ctx* c, int mode, const double* D, const double* x, const double* b, double* out,
                    const FistaState* st, double* const* rbuf, cudaStream_t s, double* lam, double rho,
                    const int* gate_nnz, int gate_cap) {
    PassParams p;
    p.lam = lam;
    p.rho = rho;
    p.gate_nnz = gate_nnz;
    p.gate_cap = gate_cap;
    p.scal = c->scal;
    p.D = D;
    p.ld = c->ld;
    p.n = c->n_local;
    p.d = c->d;
    p.x = x;
    p.b = b;
    p.out = out;
    p.gpart = c->gpart;
    p.sspart = c->sspart;
    p.st = st;
    p.r0 = rbuf ? rbuf[0] : nullptr;
    p.r1 = rbuf ? rbuf[1] : nullptr;
    p.R = c->pass_rows;
    p.stages = c->pass_stages;
    p.mode = mode;
    // stream D through L2 when it cannot stay resident (B200 L2 ~126 MB); keep it when it can
    p.evict_first = ((size_t)c->n_local * c->ld * c->esz > ((size_t)96 << 20)) ? 1 : 0;
    const size_t row_bytes = (size_t)c->ld * c->esz;
    p.stage_stride = (uint32_t)(((size_t)c->pass_rows * row_bytes + 127) & ~(size_t)127);
    const int64_t ld2 = c->ld / 2;
    const int wpr = pick_wpr(ld2);
    if (wpr == 1) {
        if (ld2 <= 128) return launch_t<4, 1, 1>(c, p, s);
        if (ld2 <= 256) return launch_t<8, 1, 1>(c, p, s);
        return launch_t<16, 1, 2>(c, p, s);
    }
    if (wpr == 2) return launch_t<16, 2, 4>(c, p, s);
    if (wpr == 4) return launch_t<16, 4, 8>(c, p, s);
    return launch_t<16, 8, 16>(c, p, s);
}

int rbl_launch_gather(rbl_ctx* c, const double* D, const int32_t* rows, const double* delta, const int* count,
                      int cap, cudaStream_t s) {
    GatherParams p;
    p.D = D;
    p.ld = c->ld;
    p.rows = rows;
    p.delta = delta;
    p.count = count;
    p.cap = cap;
    p.gpart = c->gpart;
    p.sspart = c->sspart;
    p.R = c->pass_rows > 32 ? 32 : c->pass_rows;  // one issuing lane per row
    p.stages = c->pass_stages;
    const size_t row_bytes = (size_t)c->ld * c->esz;
    p.stage_stride = (uint32_t)(((size_t)c->pass_rows * row_bytes + 127) & ~(size_t)127);
    const int64_t ld2 = c->ld / 2;
    if (ld2 <= 256) return launch_gather_t<1>(c, p, s);
    if (ld2 <= 512) return launch_gather_t<2>(c, p, s);
    if (ld2 <= 1024) return launch_gather_t<4>(c, p, s);
    if (ld2 <= 2048) return launch_gather_t<8>(c, p, s);
    return launch_gather_t<16>(c, p, s);
}
