// Host emulation of the device PAV pipeline (chunk stage + tree levels), built from the SAME
// header the CUDA kernels include (csrc/pav_core.h).  Test infrastructure: lets the CPU test-suite
// check the merge logic against the oracle's sequential stack PAV without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "pav_core.h"

static void fill(double* val, int64_t lo, int64_t hi, double v) {
    for (int64_t i = lo; i < hi; ++i) val[i] = v;
}

extern "C" int emul_pav(int loss, int64_t n, const double* sigma, const double* m, double rho, int chunk_log2,
                        double* val, int64_t* n_merges) {
    const int64_t CH = int64_t(1) << chunk_log2;
    const int64_t nch = (n + CH - 1) / CH;
    int64_t merges = 0;
    for (int64_t i = 0; i < n; ++i) val[i] = rbl_block_prox(loss, sigma[i], m[i], rho);
    // chunk-local exclusive prefixes (n entries) + chunk totals
    std::vector<double> lsh(n + 1), lsl(n + 1), lmh(n + 1), lml(n + 1);  // entry n: end of a partial last chunk
    std::vector<double> osh(nch + 1), osl(nch + 1), omh(nch + 1), oml(nch + 1);
    dd_t ts = dd_make(0.0), tm = dd_make(0.0);
    for (int64_t ch = 0; ch < nch; ++ch) {
        osh[ch] = ts.hi; osl[ch] = ts.lo; omh[ch] = tm.hi; oml[ch] = tm.lo;
        int64_t base = ch * CH, len = (n - base < CH) ? n - base : CH;
        // flat local prefix with len+1 entries for the in-chunk levels
        std::vector<double> psh(len + 1), psl(len + 1), pmh(len + 1), pml(len + 1);
        dd_t as = dd_make(0.0), am = dd_make(0.0);
        for (int64_t i = 0; i < len; ++i) {
            psh[i] = as.hi; psl[i] = as.lo; pmh[i] = am.hi; pml[i] = am.lo;
            lsh[base + i] = as.hi; lsl[base + i] = as.lo; lmh[base + i] = am.hi; lml[base + i] = am.lo;
            as = dd_add_d(as, sigma[base + i]);
            am = dd_add_d(am, m[base + i]);
        }
        psh[len] = as.hi; psl[len] = as.lo; pmh[len] = am.hi; pml[len] = am.lo;
        if (len < CH) { lsh[base + len] = as.hi; lsl[base + len] = as.lo; lmh[base + len] = am.hi; lml[base + len] = am.lo; }
        ts = dd_add(ts, as);
        tm = dd_add(tm, am);
        PrefixFlat ps{psh.data(), psl.data()}, pm{pmh.data(), pml.data()};
        double* v = val + base;
        for (int64_t w = 1; w < len; w <<= 1) {
            for (int64_t a = 0; a + w < len; a += 2 * w) {
                int64_t b = a + w, c = (a + 2 * w < len) ? a + 2 * w : len;
                int64_t lo, hi; double vv;
                bool hit = (2 * w <= 32) ? pav_merge_search(loss, rho, ValPlain{v}, ps, pm, a, b, c, &lo, &hi, &vv)
                                          : pav_merge_search_kary(loss, rho, ValPlain{v}, ps, pm, a, b, c, &lo, &hi, &vv);
                if (hit) { fill(v, lo, hi, vv); ++merges; }
            }
        }
    }
    osh[nch] = ts.hi; osl[nch] = ts.lo; omh[nch] = tm.hi; oml[nch] = tm.lo;
    PrefixChunked ps{lsh.data(), lsl.data(), osh.data(), osl.data(), chunk_log2};
    PrefixChunked pm{lmh.data(), lml.data(), omh.data(), oml.data(), chunk_log2};
    for (int64_t w = CH; w < n; w <<= 1) {
        for (int64_t a = 0; a + w < n; a += 2 * w) {
            int64_t b = a + w, c = (a + 2 * w < n) ? a + 2 * w : n;
            int64_t lo, hi; double vv;
            if (pav_merge_search_kary(loss, rho, ValPlain{val}, ps, pm, a, b, c, &lo, &hi, &vv)) { fill(val, lo, hi, vv); ++merges; }
        }
    }
    if (n_merges) *n_merges = merges;
    return 0;
}

// ---- few-segment route (csrc/pav_kernels.cu: pav_seg_merge_kernel): level-0 solved ranges are the maximal runs
// of ranks where sigma does not increase; (#runs - 1) merges of the solved prefix with the next run over a value
// array with pending pooled blocks laid over it; one fill at the end.
struct ValOverlayHost {
    const double* p;
    const std::vector<int64_t>* lo;
    const std::vector<int64_t>* hi;
    const std::vector<double>* v;
    double operator()(int64_t i) const {
        for (size_t k = 0; k < lo->size(); ++k)
            if (i >= (*lo)[k] && i < (*hi)[k]) return (*v)[k];
        return p[i];
    }
};

// hints_in (may be null): [2 (#runs - 1)] guesses (lo*, hi*) of the pooled block of every merge (-1: none), e.g. the
// blocks of an earlier call on similar margins; hints_out (may be null) receives this call's blocks the same way
extern "C" int emul_pav_fewseg_hinted(int loss, int64_t n, const double* sigma, const double* m, double rho,
                                      int chunk_log2, double* val, int64_t* n_runs, const int64_t* hints_in,
                                      int64_t* hints_out);

// probe spacing of the warm-started first round (pav_hint_pos): RBL_HINT_STRIDE, or RBL_HINT_STRIDE_NEAR as the device
// kernel uses while its guesses keep landing close — the pooled blocks never depend on it
static int g_hint_stride = RBL_HINT_STRIDE;
extern "C" void emul_set_hint_stride(int stride) { g_hint_stride = stride; }

extern "C" int emul_pav_fewseg(int loss, int64_t n, const double* sigma, const double* m, double rho, int chunk_log2,
                               double* val, int64_t* n_runs) {
    return emul_pav_fewseg_hinted(loss, n, sigma, m, rho, chunk_log2, val, n_runs, nullptr, nullptr);
}

extern "C" int emul_pav_fewseg_hinted(int loss, int64_t n, const double* sigma, const double* m, double rho,
                                      int chunk_log2, double* val, int64_t* n_runs, const int64_t* hints_in,
                                      int64_t* hints_out) {
    const int64_t CH = int64_t(1) << chunk_log2;
    const int64_t nch = (n + CH - 1) / CH;
    for (int64_t i = 0; i < n; ++i) val[i] = rbl_block_prox(loss, sigma[i], m[i], rho);
    std::vector<int64_t> bounds{0};
    for (int64_t i = 1; i < n; ++i)
        if (sigma[i] > sigma[i - 1]) bounds.push_back(i);
    bounds.push_back(n);
    if (n_runs) *n_runs = (int64_t)bounds.size() - 1;
    // chunked double-double prefixes of sigma and m, exactly as the device lays them out
    std::vector<double> lsh(n + 1), lsl(n + 1), lmh(n + 1), lml(n + 1);
    std::vector<double> osh(nch + 1), osl(nch + 1), omh(nch + 1), oml(nch + 1);
    dd_t ts = dd_make(0.0), tm = dd_make(0.0);
    for (int64_t ch = 0; ch < nch; ++ch) {
        osh[ch] = ts.hi; osl[ch] = ts.lo; omh[ch] = tm.hi; oml[ch] = tm.lo;
        int64_t base = ch * CH, len = (n - base < CH) ? n - base : CH;
        dd_t as = dd_make(0.0), am = dd_make(0.0);
        for (int64_t i = 0; i < len; ++i) {
            lsh[base + i] = as.hi; lsl[base + i] = as.lo; lmh[base + i] = am.hi; lml[base + i] = am.lo;
            as = dd_add_d(as, sigma[base + i]);
            am = dd_add_d(am, m[base + i]);
        }
        if (len < CH) { lsh[base + len] = as.hi; lsl[base + len] = as.lo; lmh[base + len] = am.hi; lml[base + len] = am.lo; }
        ts = dd_add(ts, as);
        tm = dd_add(tm, am);
    }
    osh[nch] = ts.hi; osl[nch] = ts.lo; omh[nch] = tm.hi; oml[nch] = tm.lo;
    PrefixChunked ps{lsh.data(), lsl.data(), osh.data(), osl.data(), chunk_log2};
    PrefixChunked pm{lmh.data(), lml.data(), omh.data(), oml.data(), chunk_log2};
    std::vector<int64_t> blo, bhi;
    std::vector<double> bv;
    ValOverlayHost ov{val, &blo, &bhi, &bv};
    for (size_t j = 1; j + 1 < bounds.size(); ++j) {
        const int64_t b = bounds[j], c = bounds[j + 1];
        int64_t lo, hi;
        double vv;
        const int64_t h_lo = hints_in ? hints_in[2 * (j - 1)] : -1, h_hi = hints_in ? hints_in[2 * (j - 1) + 1] : -1;
        if (hints_out) hints_out[2 * (j - 1)] = hints_out[2 * (j - 1) + 1] = -1;
        if (!pav_merge_search_kary(loss, rho, ov, ps, pm, (int64_t)0, b, c, &lo, &hi, &vv, h_lo, h_hi, g_hint_stride)) continue;
        if (hints_out) {
            hints_out[2 * (j - 1)] = lo;
            hints_out[2 * (j - 1) + 1] = hi;
        }
        std::vector<int64_t> nlo, nhi;
        std::vector<double> nv;
        for (size_t k = 0; k < blo.size(); ++k) {
            if (blo[k] >= lo && bhi[k] <= hi) continue;  // swallowed whole
            nlo.push_back(blo[k]); nhi.push_back(bhi[k]); nv.push_back(bv[k]);
        }
        nlo.push_back(lo); nhi.push_back(hi); nv.push_back(vv);
        blo = nlo; bhi = nhi; bv = nv;
    }
    for (size_t k = 0; k < blo.size(); ++k) fill(val, blo[k], bhi[k], bv[k]);
    return 0;
}

extern "C" uint64_t emul_key(double x) { uint64_t b; memcpy(&b, &x, 8); return rbl_key_from_bits(b); }
extern "C" double emul_unkey(uint64_t k) { uint64_t b = rbl_bits_from_key(k); double x; memcpy(&x, &b, 8); return x; }
extern "C" double emul_prox(int loss, double s, double m, double rho) { return rbl_block_prox(loss, s, m, rho); }

// ---- the library's L-BFGS-B (csrc/lbfgs_core.h) driven by a host callback, for comparison with scipy -----------
#include "lbfgs_core.h"
typedef int (*emul_fg_t)(const double* x, double* f, double* g);
extern "C" int emul_lbfgs(int n, double* x, emul_fg_t fg, int m, int maxiter, int* nit, int* nfev, double* f_out) {
    rbl_lbfgs::Result r = rbl_lbfgs::minimize(n, x, [&](const double* xx, double* f, double* g) { return fg(xx, f, g); },
                                              m, maxiter);
    *nit = r.nit;
    *nfev = r.nfev;
    *f_out = r.f;
    return r.status;
}
