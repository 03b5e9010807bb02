// metrics_kernels.cu — test-set metrics: the step after the hot path.
//
// Replaces (semantics): src/util/calculate_acc.py:3-19 (accuracy) and src/util/fair_metric.py:3-40 (group
// confusion counts, Theil index sums).  The reference forms X_test @ w, a probability vector, several boolean
// masks and a dozen fancy-indexed temporaries; here ONE pass over X_test produces the 16 numbers all of those
// results are functions of.  HBM-bound like every other D pass: n*d*8 algorithmic bytes, X read exactly once.
//
// A warp takes a row: lanes stride the columns 16 bytes at a time (rows staged through a cp.async ring in shared
// memory, w in shared memory too), a butterfly adds the 32 partial dots in a fixed order; after 32 rows every lane classifies one of them (exp / log 32 wide) into
// per-lane counters.  Counts are exact integers; the two floating-point sums are reduced in a fixed order (lanes ->
// warp -> CTA -> last CTA over the per-CTA partials), so a call is reproducible run to run.
#include "common.cuh"

namespace {

constexpr int kMT = 256;    // threads per CTA: 8 rows in flight per CTA
constexpr int kMRing = 2048;  // doubles in a warp's row ring (16 KB)
constexpr int kMS = 4;        // units in the ring (kMS - 1 in flight); a power of two
constexpr int kMJ = 8;        // 16-byte pieces per lane per unit: kMRing * 8 == kMS * kMJ * 32 * 16
constexpr int kMOut = 16;   // numbers per call (layout: include/rbl_b200.h, rbl_test_metrics)

struct MetricsParams {
    const double* X;
    int64_t n, d, ld;
    const double* w;
    const double* y;
    const int32_t* group;   // may be null: every row in group 0
    int loss;
    int w_shared;           // w staged in shared memory (else read through L1/L2)
    int vec2;               // rows are 16-byte aligned and d >= 2: double2 loads
    double threshold;
    double* partial;        // [gridDim.x][kMOut]
    unsigned int* ticket;   // zero on entry, zero again on exit
    double* out;            // [kMOut]
};

template <int LPR>  // lanes per row: 32 for d > 256, fewer for narrow rows
__global__ void __launch_bounds__(kMT, 1) metrics_kernel(const MetricsParams p) {
    static_assert(kMRing * 8 == kMS * kMJ * 32 * 16 && (kMS & (kMS - 1)) == 0, "ring geometry");
    extern __shared__ __align__(16) double sring[];  // [8 warps][kMRing] row rings, then w [d] when w_shared
    double* sw = sring + (kMT / 32) * kMRing;
    __shared__ double s_acc[kMT / kMOut][kMOut];  // rows 0..7: the warps; all 16: shares of the last CTA
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* w = p.w;
    if (p.w_shared) {
        for (int64_t i = tid; i < p.d; i += kMT) sw[i] = p.w[i];
        __syncthreads();
        w = sw;
    }
    // per-lane counters: the integer ones fit 32 bits (a warp sees n / (8 gridDim.x) rows), the two sums are doubles
    unsigned int cnt[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) cnt[k] = 0u;
    double sum_b = 0.0, sum_bl = 0.0;
    // lane j classifies the j-th row of every batch of 32 rows its warp has reduced: exp / log run 32 wide
    auto classify = [&](double s, int64_t row) {
        // calculate_acc.py:5-8 / fair_metric.py:4-7: the logistic function on the stable side of 0
        const double e = exp(-fabs(s));
        const double prob = s >= 0.0 ? 1.0 / (1.0 + e) : e / (e + 1.0);
        const bool pred = prob >= p.threshold;
        const double y = p.y[row];
        // calculate_acc.py:9-11 (+1 / -1 against the label); :13-16 hinge: the zeros of (x.w >= 0) are rewritten
        // to 1, so every prediction is +1 — reproduced as shipped
        const bool hit = (p.loss == RBL_LOSS_HINGE) ? (y == 1.0) : (y == (pred ? 1.0 : -1.0));
        cnt[0] += hit ? 1u : 0u;
        const double y01 = (y == -1.0) ? 0.0 : y;  // fair_metric.py:9-10
        const int g = p.group ? p.group[row] : 0;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {           // :11-24; rows of other groups only enter the Theil index
            const unsigned int in = (g == gg) ? 1u : 0u;
            cnt[2 + 6 * gg + 0] += in;
            cnt[2 + 6 * gg + 1] += pred ? in : 0u;
            cnt[2 + 6 * gg + 2] += (pred && y01 == 1.0) ? in : 0u;    // TP
            cnt[2 + 6 * gg + 3] += (!pred && y01 != 0.0) ? in : 0u;   // FN
            cnt[2 + 6 * gg + 4] += (!pred && y01 == 0.0) ? in : 0u;   // TN
            cnt[2 + 6 * gg + 5] += (pred && y01 != 1.0) ? in : 0u;    // FP
        }
        const double b = prob - y01 + 1.0;         // :35
        sum_b += b;
        sum_bl += b * log(b);                      // b == 0: 0 * -inf = nan, as numpy gives
    };
    const int64_t nwarps = (int64_t)gridDim.x * (kMT / 32);
    const int64_t g0 = (int64_t)blockIdx.x * (kMT / 32) + warp;
    double my_s = 0.0;
    int64_t my_row = -1;
    if (p.vec2) {
        // LPR lanes share a row, a warp step covers RPW = 32 / LPR consecutive rows (narrow rows: several per step,
        // so the per-row reduction and the bytes in flight do not degrade with d).  Rows arrive through a per-warp
        // ring of kMS units of 4 KB: every lane copies its own 8 x 16 bytes of a unit with cp.async (no registers
        // held while in flight: 3 units = 12 KB per warp, 96 KB per SM) and later reads back exactly those bytes,
        // so no cross-lane synchronisation is needed.
        constexpr int RPW = 32 / LPR, CPU = 2 * LPR * kMJ;  // CPU: columns of a row per unit
        const int sub = lane % LPR, rsel = lane / LPR;
        const int64_t dv = p.d & ~(int64_t)1;
        const int nck = (int)((dv + CPU - 1) / CPU);
        const int64_t ngroups = (p.n + RPW - 1) / RPW;
        unsigned char* ring = reinterpret_cast<unsigned char*>(sring) + (size_t)warp * (kMRing * 8);
        const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16;
        // issue / consume cursors advance incrementally (row group, chunk, ring slot): no divisions in the loop
        int64_t i_g = g0;
        int i_k = 0, i_slot = 0;
        auto issue = [&]() {
            if (i_g < ngroups) {
                const int64_t row = i_g * RPW + rsel;
                const int cbase = i_k * CPU + 2 * sub;
                if (row < p.n) {
                    const double* src = p.X + row * p.ld + cbase;
                    const uint32_t dst = ring_s + (uint32_t)i_slot * (kMRing * 8 / kMS);
#pragma unroll
                    for (int j = 0; j < kMJ; ++j)
                        if (cbase + 2 * LPR * j < dv)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 512),
                                         "l"(src + 2 * LPR * j)
                                         : "memory");
                }
                if (++i_k == nck) {
                    i_k = 0;
                    i_g += nwarps;
                }
                i_slot = (i_slot + 1) & (kMS - 1);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int u = 0; u < kMS - 1; ++u) issue();
        double s0 = 0.0, s1 = 0.0;
        int c_k = 0, c_slot = 0, it = 0;
        for (int64_t g = g0; g < ngroups;) {
            issue();
            asm volatile("cp.async.wait_group %0;" ::"n"(kMS - 1) : "memory");
            const int64_t row = g * RPW + rsel;
            const unsigned char* src = ring + (size_t)c_slot * (kMRing * 8 / kMS) + lane * 16;
            const int cbase = c_k * CPU + 2 * sub;
            if (row < p.n) {
#pragma unroll
                for (int j = 0; j < kMJ; ++j) {
                    const int c = cbase + 2 * LPR * j;
                    if (c < dv) {
                        const double2 v = *reinterpret_cast<const double2*>(src + j * 512);
                        const double2 ww = *reinterpret_cast<const double2*>(w + c);
                        s0 = fma(v.x, ww.x, s0);
                        s1 = fma(v.y, ww.y, s1);
                    }
                }
            }
            c_slot = (c_slot + 1) & (kMS - 1);
            if (++c_k == nck) {
                if (dv != p.d && sub == 0 && row < p.n) s0 = fma(p.X[row * p.ld + dv], w[dv], s0);
                double sr = s0 + s1;
#pragma unroll
                for (int o = LPR / 2; o; o >>= 1) sr += __shfl_xor_sync(0xffffffffu, sr, o);
                // lane l keeps the row of its sub-group at step it = l mod LPR: after LPR steps every lane holds one
                // row, and exp / log run 32 wide
                if (sub == (it % LPR)) {
                    my_s = sr;
                    my_row = row < p.n ? row : -1;
                }
                if ((++it % LPR) == 0) {
                    if (my_row >= 0) classify(my_s, my_row);
                    my_row = -1;
                }
                s0 = s1 = 0.0;
                c_k = 0;
                g += nwarps;
            }
        }
    } else {
        // rows that are not 16-byte aligned (direct callers only; DeviceTestSet pads): plain warp-per-row loads
        for (int64_t row = g0; row < p.n; row += nwarps) {
            const double* x = p.X + row * p.ld;
            double sr = 0.0;
#pragma unroll 4
            for (int64_t c = lane; c < p.d; c += 32) sr = fma(__ldg(x + c), w[c], sr);
#pragma unroll
            for (int o = 16; o; o >>= 1) sr += __shfl_xor_sync(0xffffffffu, sr, o);
            if (lane == 0) classify(sr, row);
        }
    }
    if (my_row >= 0) classify(my_s, my_row);
    // lanes -> warp (butterfly: a fixed order), then the 16 numbers of the warp as doubles
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 14; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
        sum_b += __shfl_xor_sync(0xffffffffu, sum_b, o);
        sum_bl += __shfl_xor_sync(0xffffffffu, sum_bl, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 14; ++k) s_acc[warp][k] = (double)cnt[k];
        s_acc[warp][14] = sum_b;
        s_acc[warp][15] = sum_bl;
    }
    __syncthreads();
    if (tid < kMOut) {
        double t = 0.0;
        for (int wq = 0; wq < kMT / 32; ++wq) t += s_acc[wq][tid];
        p.partial[(size_t)blockIdx.x * kMOut + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: 16 threads per output, each a fixed strided share of the CTAs, then a fixed-order finish
    {
        const int col = tid & (kMOut - 1), part = tid / kMOut;  // kMT / kMOut = 16 parts
        double t = 0.0;
        for (unsigned int b = part; b < gridDim.x; b += kMT / kMOut) t += __ldcg(p.partial + (size_t)b * kMOut + col);
        s_acc[part][col] = t;  // (every read of the per-warp rows is behind the barriers above)
    }
    __syncthreads();
    if (tid < kMOut) {
        double t = 0.0;
        for (int part = 0; part < kMT / kMOut; ++part) t += s_acc[part][tid];
        p.out[tid] = (tid == 1) ? (double)p.n : t;
    }
    if (tid == 0) *p.ticket = 0u;
}

#define RBL_TRY_K(x)                \
    do {                            \
        const int rc__ = (x);       \
        if (rc__ != RBL_OK) return rc__; \
    } while (0)

}  // namespace

size_t rbl_k_metrics_scratch_bytes(int num_sms) { return ((size_t)num_sms * 4 * kMOut + 2) * sizeof(double); }

int rbl_k_test_metrics(int num_sms, const double* X, int64_t n, int64_t d, int64_t ld, const double* w,
                       const double* y, const int32_t* group, int loss, double threshold, double* out16,
                       void* scratch, cudaStream_t s) {
    MetricsParams p;
    p.X = X;
    p.n = n;
    p.d = d;
    p.ld = ld;
    p.w = w;
    p.y = y;
    p.group = group;
    p.loss = loss;
    p.threshold = threshold;
    int grid = num_sms;  // persistent: one CTA per SM (128 KB of row ring each)
    const int64_t need = (n + kMT / 32 - 1) / (kMT / 32);
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    p.partial = reinterpret_cast<double*>(scratch);
    p.ticket = reinterpret_cast<unsigned int*>(p.partial + (size_t)num_sms * 4 * kMOut);
    p.out = out16;
    const size_t ring = (size_t)(kMT / 32) * kMRing * sizeof(double);  // 128 KB
    const size_t wbytes = (size_t)d * sizeof(double);
    p.w_shared = wbytes <= 64 * 1024 ? 1 : 0;  // wider w comes through L1 / L2
    p.vec2 = (d >= 2 && (ld % 2) == 0 && (reinterpret_cast<uintptr_t>(X) % 16) == 0 &&
              (p.w_shared || (reinterpret_cast<uintptr_t>(w) % 16) == 0)) ? 1 : 0;
    const size_t smem = ring + (p.w_shared ? wbytes : 0);
    auto launch = [&](auto kernel) -> int {
        // (the attribute is per device, so it is set on every call: microseconds against a pass of hundreds)
        RBL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ring + 64 * 1024)));
        kernel<<<grid, kMT, smem, s>>>(p);
        return RBL_OK;
    };
    const int64_t per_lane = 2 * kMJ;  // columns a lane covers per unit
    if (!p.vec2 || d > 16 * per_lane) RBL_TRY_K(launch(metrics_kernel<32>));
    else if (d > 8 * per_lane) RBL_TRY_K(launch(metrics_kernel<16>));
    else if (d > 4 * per_lane) RBL_TRY_K(launch(metrics_kernel<8>));
    else if (d > 2 * per_lane) RBL_TRY_K(launch(metrics_kernel<4>));
    else if (d > per_lane) RBL_TRY_K(launch(metrics_kernel<2>));
    else RBL_TRY_K(launch(metrics_kernel<1>));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}
