// ingest_kernels.cu — the step before the hot path: column standardisation of the raw
// feature matrix and the row gather of the train / test split, on the device.
//
// Replaces (semantics): src/util/load_data.py:115 `preprocessing.scale(X)` (scikit-learn: per column
// x <- (x - mean) / std, std with ddof = 0, columns with std < 10 eps keep scale 1) and the
// `train_test_split` row selection of the drivers (run_SRM.py:26): X[idx].  Both are HBM-bound streaming passes
// over the n x d matrix: standardisation reads X three times and writes it once (mean; centred sum of squares —
// the two-pass form numpy's std uses; transform), the split reads and writes every selected row once.
#include "common.cuh"

namespace {

constexpr int kIT = 256;   // threads per CTA: one double2 column pair each -> 512 columns per CTA
constexpr int kIU = 8;     // rows in flight per thread

// partial[strip][col] = sum over the rows of the strip of x (mode 0) or (x - mean)^2 (mode 1)
__global__ void __launch_bounds__(kIT) colsum_kernel(const double* __restrict__ X, int64_t n, int64_t ld2,
                                                     int64_t rows_per_strip, int mode,
                                                     const double* __restrict__ mean, double* __restrict__ partial) {
    const int64_t c2 = (int64_t)blockIdx.y * kIT + threadIdx.x;  // column pair
    if (c2 >= ld2) return;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_strip;
    const int64_t r1 = r0 + rows_per_strip < n ? r0 + rows_per_strip : n;
    double2 mu = make_double2(0.0, 0.0);
    if (mode) mu = reinterpret_cast<const double2*>(mean)[c2];
    const double2* x = reinterpret_cast<const double2*>(X) + c2;
    // kIU independent accumulator pairs: no serial chain behind the kIU loads of a round (they stay in flight
    // together); combined in a fixed order at the end
    double ax[kIU], ay[kIU];
#pragma unroll
    for (int u = 0; u < kIU; ++u) ax[u] = ay[u] = 0.0;
    int64_t r = r0;
    for (; r + kIU <= r1; r += kIU) {
        double2 v[kIU];
#pragma unroll
        for (int u = 0; u < kIU; ++u) v[u] = __ldcs(x + (r + u) * ld2);
#pragma unroll
        for (int u = 0; u < kIU; ++u) {
            const double a = v[u].x - mu.x, b = v[u].y - mu.y;
            ax[u] += mode ? a * a : a;
            ay[u] += mode ? b * b : b;
        }
    }
    for (; r < r1; ++r) {
        const double2 v = __ldcs(x + r * ld2);
        const double a = v.x - mu.x, b = v.y - mu.y;
        ax[0] += mode ? a * a : a;
        ay[0] += mode ? b * b : b;
    }
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int u = 0; u < kIU; ++u) {
        sx += ax[u];
        sy += ay[u];
    }
    reinterpret_cast<double2*>(partial)[(int64_t)blockIdx.x * ld2 + c2] = make_double2(sx, sy);
}

// strips summed in a fixed order; mode 0: mean = sum / n; mode 1: scale = sqrt(sum / n), 1 where < 10 eps
__global__ void colsum_finish_kernel(const double* __restrict__ partial, int nstrips, int64_t ld, int64_t n,
                                     int mode, double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ld) return;
    double s = 0.0;
    for (int k = 0; k < nstrips; ++k) s += partial[(int64_t)k * ld + c];
    if (mode == 0) {
        out[c] = s / (double)n;
    } else {
        const double sd = sqrt(s / (double)n);
        out[c] = sd < 10.0 * 2.220446049250313e-16 ? 1.0 : sd;  // sklearn _handle_zeros_in_scale
    }
}

__global__ void __launch_bounds__(kIT) standardize_kernel(double* __restrict__ X, int64_t n, int64_t ld2,
                                                          const double* __restrict__ mean,
                                                          const double* __restrict__ scale) {
    const int64_t c2 = (int64_t)blockIdx.y * kIT + threadIdx.x;
    if (c2 >= ld2) return;
    const double2 mu = reinterpret_cast<const double2*>(mean)[c2], sc = reinterpret_cast<const double2*>(scale)[c2];
    double2* x = reinterpret_cast<double2*>(X) + c2;
    for (int64_t r0 = (int64_t)blockIdx.x * kIU; r0 < n; r0 += (int64_t)gridDim.x * kIU) {
        double2 v[kIU];
#pragma unroll
        for (int u = 0; u < kIU; ++u)
            if (r0 + u < n) v[u] = __ldcs(x + (r0 + u) * ld2);
#pragma unroll
        for (int u = 0; u < kIU; ++u)
            if (r0 + u < n) __stcs(x + (r0 + u) * ld2, make_double2((v[u].x - mu.x) / sc.x, (v[u].y - mu.y) / sc.y));
    }
}

// out[i, :] = X[idx[i], :]: a warp per row, 16-byte pieces
__global__ void __launch_bounds__(kIT) gather_rows_kernel(const double* __restrict__ X, int64_t ld2_in,
                                                          const int64_t* __restrict__ idx, int64_t n_out,
                                                          int64_t ld2_out, int64_t d2, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (kIT / 32);
    for (int64_t i = (int64_t)blockIdx.x * (kIT / 32) + (threadIdx.x >> 5); i < n_out; i += nwarps) {
        const double2* src = reinterpret_cast<const double2*>(X) + idx[i] * ld2_in;
        double2* dst = reinterpret_cast<double2*>(out) + i * ld2_out;
        for (int64_t c = lane; c < d2; c += 32) __stcs(dst + c, __ldcs(src + c));
    }
}

}  // namespace

int rbl_k_standardize_scratch_doubles(int num_sms, int64_t ld, int64_t* out) {
    const int64_t ld2 = ld / 2, ytiles = (ld2 + kIT - 1) / kIT;
    int64_t strips = ((int64_t)num_sms * 8 + ytiles - 1) / ytiles;
    if (strips < 1) strips = 1;
    *out = strips * ld;
    return (int)strips;
}

int rbl_k_standardize(int num_sms, double* X, int64_t n, int64_t ld, double* mean, double* scale, double* scratch,
                      cudaStream_t s) {
    const int64_t ld2 = ld / 2, ytiles = (ld2 + kIT - 1) / kIT;
    int64_t cap = 0;
    int strips = rbl_k_standardize_scratch_doubles(num_sms, ld, &cap);
    if ((int64_t)strips > n) strips = (int)n;
    const int64_t rps = (n + strips - 1) / strips;
    strips = (int)((n + rps - 1) / rps);
    const dim3 grid(strips, (unsigned)ytiles);
    const int fin_grid = (int)((ld + 255) / 256);
    colsum_kernel<<<grid, kIT, 0, s>>>(X, n, ld2, rps, 0, nullptr, scratch);
    RBL_LAUNCH_CHECK();
    colsum_finish_kernel<<<fin_grid, 256, 0, s>>>(scratch, strips, ld, n, 0, mean);
    RBL_LAUNCH_CHECK();
    colsum_kernel<<<grid, kIT, 0, s>>>(X, n, ld2, rps, 1, mean, scratch);
    RBL_LAUNCH_CHECK();
    colsum_finish_kernel<<<fin_grid, 256, 0, s>>>(scratch, strips, ld, n, 1, scale);
    RBL_LAUNCH_CHECK();
    int xg = (int)((n + kIU - 1) / kIU);
    const int xcap = (int)(((int64_t)num_sms * 8 + ytiles - 1) / ytiles);
    if (xg > xcap) xg = xcap;
    standardize_kernel<<<dim3(xg < 1 ? 1 : xg, (unsigned)ytiles), kIT, 0, s>>>(X, n, ld2, mean, scale);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_k_gather_rows(int num_sms, const double* X, int64_t ld_in, const int64_t* idx, int64_t n_out, int64_t d,
                      double* out, int64_t ld_out, cudaStream_t s) {
    int64_t grid = (n_out + kIT / 32 - 1) / (kIT / 32);
    if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
    gather_rows_kernel<<<(int)grid, kIT, 0, s>>>(X, ld_in / 2, idx, n_out, ld_out / 2, (d + 1) / 2, out);
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}
