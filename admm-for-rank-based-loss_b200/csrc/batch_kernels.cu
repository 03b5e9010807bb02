// batch_kernels.cu — batched mode (K10): many independent ADMM instances (lambda grid, seeds) that
// share ONE design matrix D.  No reference counterpart (the reference solves one instance per
// Python object; the oracle is a loop over its ADMMmethod): SURVEY.md §2.2 K10, §8e "instance sharding".
//
// The pass over D is the same TMA-staged row-tile pipeline as pass_kernels.cu, but every tile is
// applied to G = 8 right-hand sides at once:
//     R[8 rows x 8 inst]   = B - Dtile[8 x d] . X[d x 8]        (phase A)
//     Gacc[d x 8 inst]    += Dtile^T[d x 8 rows] . R[8 x 8]     (phase B)
// i.e. two tall-skinny fp64 GEMMs per tile.  With 8 instances the arithmetic intensity is 8x the
// single-instance matvec (2 flop/byte): about as much FP64 FMA time as HBM time on B200, so the
// products run on the FP64 tensor-core path (mma.sync m8n8k4 f64 — there is no tcgen05 kind for
// fp64) with both operands read once per MMA from shared memory.  Rows are staged with a padded
// leading dimension (ldp = 4 mod 16 doubles) so the 8x4 / 4x8 fragment loads are bank-conflict free.
#include "common.cuh"

namespace {

constexpr int kG = 8;         // instances per pass
constexpr int kRows = 8;      // rows per tile (= MMA M)
constexpr int kBThreads = 256;
constexpr int kBWarps = 8;
constexpr int kMaxGroupsPerWarp = 16;  // column groups of 8 per warp: d <= 8*8*16 = 1024

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
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor core
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct MultiParams {
    const double* D;
    int64_t ld;
    int64_t n;
    int d;
    const double* x;       // [G][xstride] trial points of the group
    int64_t xstride;
    const double* b;       // [G][n]
    double* r0;            // [G][n] residual buffers (per instance the one FistaState.cur selects)
    double* r1;
    const FistaState* st;  // [G] (nullptr: always write r0)
    double* gpart;         // [grid][G][ld]
    double* sspart;        // [grid][G]
    int ldp;               // padded smem leading dimension (doubles), = 4 mod 16
    int stages;
    int ninst;             // valid instances in this group (<= G)
};

__global__ void __launch_bounds__(kBThreads, 1) rbl_pass_multi_kernel(const MultiParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldp = p.ldp, S = p.stages;
    const size_t stage_doubles = (size_t)kRows * ldp;
    double* stage0 = reinterpret_cast<double*>(smem);
    double* xs = stage0 + (size_t)S * stage_doubles;          // [G][ldp]
    double* cpart = xs + (size_t)kG * ldp;                     // [warps][8][8]
    double* rt = cpart + kBWarps * 64;                         // [8 rows][8 inst]
    double* ssred = rt + 64;                                   // [8][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ssred + 64);  // [S]

    if (p.st) {  // every instance of the group already converged: nothing to do
        bool all_done = true;
        for (int g = 0; g < p.ninst; ++g) all_done = all_done && (p.st[g].done != 0);
        if (all_done) return;
    }
    const int64_t ntiles = (p.n + kRows - 1) / kRows;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t nmine = ntiles > first ? (ntiles - first + stride - 1) / stride : 0;

    // zero the staging ring (padding columns and never-written rows must be finite) and load X
    for (size_t i = tid; i < (size_t)S * stage_doubles; i += kBThreads) stage0[i] = 0.0;
    for (int i = tid; i < kG * ldp; i += kBThreads) {
        const int g = i / ldp, c = i - g * ldp;
        xs[i] = (g < p.ninst && c < p.d) ? p.x[(size_t)g * p.xstride + c] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    // make the generic-proxy zero fill visible to the async proxy before TMA writes land on top of it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    const uint32_t row_bytes = (uint32_t)(p.ld * sizeof(double));
    auto issue = [&](int64_t k) {
        const int s = (int)(k % S);
        const int64_t row0 = (first + k * stride) * kRows;
        const int rows = (int)((p.n - row0 < kRows) ? (p.n - row0) : kRows);
        mbar_expect_tx(&bars[s], rows * row_bytes);
        double* dst = stage0 + (size_t)s * stage_doubles;
        for (int r = 0; r < rows; ++r) tma_bulk_g2s(dst + (size_t)r * ldp, p.D + (row0 + r) * p.ld, row_bytes, &bars[s]);
    };
    if (tid == 0) {
        const int64_t pre = nmine < S ? nmine : S;
        for (int64_t k = 0; k < pre; ++k) issue(k);
    }

    // residual buffer per instance
    double* rout = nullptr;
    const int my_row = tid >> 3, my_g = tid & 7;  // threads 0..63 own one (row, instance) of the tile
    if (tid < 64 && my_g < p.ninst) rout = ((p.st && p.st[my_g].cur) ? p.r1 : p.r0) + (size_t)my_g * p.n;

    const int nquads = ldp >> 2;               // k-steps of phase A
    const int ngroups = (int)((p.ld + 7) >> 3);  // 8-column groups of phase B
    double acc[kMaxGroupsPerWarp][2];
#pragma unroll
    for (int i = 0; i < kMaxGroupsPerWarp; ++i) acc[i][0] = acc[i][1] = 0.0;
    double ss = 0.0;
    const int fr = lane >> 2, fk = lane & 3;   // fragment coordinates

    for (int64_t k = 0; k < nmine; ++k) {
        const int s = (int)(k % S);
        const uint32_t parity = (uint32_t)((k / S) & 1);
        const int64_t row0 = (first + k * stride) * kRows;
        const int rows = (int)((p.n - row0 < kRows) ? (p.n - row0) : kRows);
        double bv = 0.0;
        if (rout && my_row < rows) bv = p.b[(size_t)my_g * p.n + row0 + my_row];
        while (!mbar_try_wait(&bars[s], parity)) {
        }
        const double* T = stage0 + (size_t)s * stage_doubles;

        // ---- phase A: C[row][inst] = sum_k T[row][k] X[k][inst]; warps split k, then reduce
        double c0 = 0.0, c1 = 0.0;
        for (int q = warp; q < nquads; q += kBWarps) {
            const int k0 = q << 2;
            const double a = T[(size_t)fr * ldp + k0 + fk];   // A[row=fr][k=fk]
            const double bb = xs[(size_t)fr * ldp + k0 + fk];  // B[k=fk][n=fr] = X[k][inst=fr]
            dmma(c0, c1, a, bb);
        }
        cpart[warp * 64 + fr * 8 + fk * 2] = c0;      // C[row=fr][inst=2fk], [2fk+1]
        cpart[warp * 64 + fr * 8 + fk * 2 + 1] = c1;
        __syncthreads();
        if (tid < 64) {
            double dot = 0.0;
#pragma unroll
            for (int w = 0; w < kBWarps; ++w) dot += cpart[w * 64 + tid];
            double r = 0.0;
            if (rout && my_row < rows) {
                r = bv - dot;
                rout[row0 + my_row] = r;
                ss = fma(r, r, ss);
            }
            rt[tid] = r;  // rt[row][inst]; zero for rows beyond the tile / unused instances
        }
        __syncthreads();
        // ---- phase B: G[col][inst] += sum_row T[row][col] R[row][inst]; each warp owns column groups
        const double rb0 = rt[fk * 8 + fr];        // B[k=row fk][n=inst fr]
        const double rb1 = rt[(fk + 4) * 8 + fr];
#pragma unroll
        for (int i = 0; i < kMaxGroupsPerWarp; ++i) {
            const int jg = warp + i * kBWarps;
            if (jg < ngroups) {
                const int cbase = jg << 3;
                const double a_lo = T[(size_t)fk * ldp + cbase + fr];        // A[m=col fr][k=row fk]
                const double a_hi = T[(size_t)(fk + 4) * ldp + cbase + fr];
                dmma(acc[i][0], acc[i][1], a_lo, rb0);
                dmma(acc[i][0], acc[i][1], a_hi, rb1);
            }
        }
        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && k + S < nmine) issue(k + S);
    }

    // ---- epilogue: per-CTA partials gpart[block][inst][col], sspart[block][inst]
    double* gp = p.gpart + (size_t)blockIdx.x * kG * p.ld;
#pragma unroll
    for (int i = 0; i < kMaxGroupsPerWarp; ++i) {
        const int jg = warp + i * kBWarps;
        if (jg < ngroups) {
            const int col = (jg << 3) + fr;  // C[m=col fr][n=inst 2fk, 2fk+1]
            if (col < p.ld) {
                gp[(size_t)(2 * fk) * p.ld + col] = acc[i][0];
                gp[(size_t)(2 * fk + 1) * p.ld + col] = acc[i][1];
            }
        }
    }
    if (tid < 64) ssred[tid] = ss;  // [row][inst]
    __syncthreads();
    if (tid < kG) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += ssred[r * 8 + tid];
        p.sspart[(size_t)blockIdx.x * kG + tid] = t;
    }
}

// red[inst][0..d) = sum over CTAs of gpart, red[inst][d] = ||r||^2, red[inst][d+1] = c0  (fixed order)
__global__ void __launch_bounds__(256) reduce_multi_kernel(const double* __restrict__ gpart,
                                                           const double* __restrict__ sspart,
                                                           const double* __restrict__ c0part, int nparts, int nc0,
                                                           int64_t ld, int d, double* __restrict__ red,
                                                           int64_t red_stride, const FistaState* st, int ninst) {
    __shared__ double sh[8][33];
    const int g = blockIdx.y;
    if (g >= ninst || (st && st[g].done)) return;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    double a = 0.0;
    if (c < d)
        for (int k = ry; k < nparts; k += 8) a += gpart[((size_t)k * kG + g) * ld + c];
    sh[ry][cx] = a;
    __syncthreads();
    double* out = red + (size_t)g * red_stride;
    if (ry == 0 && c < d) {
        double t = sh[0][cx];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += sh[q][cx];
        out[c] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double s = 0.0;
        for (int k = lane; k < nparts; k += 32) s += sspart[(size_t)k * kG + g];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        double c0 = 0.0;
        for (int k = lane; k < nc0; k += 32) c0 += c0part[(size_t)g * nc0 + k];
        for (int o = 16; o; o >>= 1) c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        if (lane == 0) {
            out[d] = s;
            out[d + 1] = c0;
        }
    }
}

}  // namespace

int rbl_batch_ldp(int64_t ld) {
    int64_t v = ((ld + 7) / 8) * 8;
    while (v % 16 != 4) v += 4;
    return (int)v;
}

size_t rbl_batch_smem(int64_t ld, int stages) {
    const int ldp = rbl_batch_ldp(ld);
    return ((size_t)stages * kRows * ldp + (size_t)kG * ldp + kBWarps * 64 + 64 + 64) * sizeof(double) +
           8 * sizeof(uint64_t);
}

int rbl_batch_group() { return kG; }

// one multi-RHS pass for the group of `ninst` (<= 8) instances starting at x / b / r / st
int rbl_k_pass_multi(rbl_ctx* c, const double* D, const double* x, int64_t xstride, const double* b, double* r0,
                     double* r1, const FistaState* st, int ninst, double* red, int64_t red_stride, const double* c0part,
                     cudaStream_t s) {
    if (c->ld > 8 * kBWarps * kMaxGroupsPerWarp) {
        rbl_set_error("batched mode supports d <= %d (got ld = %lld)", 8 * kBWarps * kMaxGroupsPerWarp,
                      (long long)c->ld);
        return RBL_ERR_UNSUPPORTED;
    }
    MultiParams p;
    p.D = D;
    p.ld = c->ld;
    p.n = c->n_local;
    p.d = c->d;
    p.x = x;
    p.xstride = xstride;
    p.b = b;
    p.r0 = r0;
    p.r1 = r1;
    p.st = st;
    p.gpart = c->bgpart;
    p.sspart = c->bsspart;
    p.ldp = rbl_batch_ldp(c->ld);
    p.stages = c->batch_stages;
    p.ninst = ninst;
    const size_t smem = rbl_batch_smem(c->ld, c->batch_stages);
    RBL_PER_DEVICE(size_t, attr, c);
    if (smem > attr) {
        RBL_CUDA(cudaFuncSetAttribute(rbl_pass_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    rbl_pass_multi_kernel<<<c->pass_grid, kBThreads, smem, s>>>(p);
    RBL_LAUNCH_CHECK();
    dim3 grid((c->d + 31) / 32, ninst);
    reduce_multi_kernel<<<grid, 256, 0, s>>>(c->bgpart, c->bsspart, c0part, c->pass_grid, c->vec_grid, c->ld, c->d,
                                             red, red_stride, st, ninst);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}
