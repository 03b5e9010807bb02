// api.cu — the C ABI declared in include/rbl_b200.h: handle, scratch, and thin launch wrappers.
#include <math.h>
#include <stdarg.h>

#include <thread>
#include <vector>
#include <stdlib.h>
#include <string.h>

#include "../../include/rbl_b200.h"
#include "common.cuh"
#include "lbfgs_core.h"

// launchers defined in the kernel translation units
int rbl_k_build_design(rbl_ctx* c, const double* X, int64_t ldx, const double* y, double* D, int64_t nrows,
                       cudaStream_t s);
int rbl_k_reduce_partials(rbl_ctx* c, int with_c0, const FistaState* st, cudaStream_t s);
int rbl_k_fista_update(rbl_ctx* c, cudaStream_t s);
int rbl_k_fista_result(rbl_ctx* c, double* w_out, double* r_out, cudaStream_t s);
int rbl_k_margins(rbl_ctx* c, const double* Dw, const double* lam, double rho, double* m, cudaStream_t s);
int rbl_k_scatter(rbl_ctx* c, const double* zs, const int32_t* perm, int use_clip, double clip, const double* lam,
                  double rho, double* z, double* b, cudaStream_t s);
int rbl_k_dual(rbl_ctx* c, const double* z, double* Dw, const double* b, const double* r, int from_residual,
               double* lam, double rho, const double* w, const double* w_prev, double* out4, cudaStream_t s);
int rbl_k_scatter_active(rbl_ctx* c, const double* zs, const double* ms, const int32_t* perm, int use_clip,
                         double clip, const double* lam, double rho, double* z, double* b, cudaStream_t s);
int rbl_k_ehrm_sums(rbl_ctx* c, const double* ms, const double* sa, const double* sb, double B, double rho,
                    double* out2, cudaStream_t s);
int rbl_k_objective(rbl_ctx* c, const double* u_sorted, const double* sigma, int loss, const double* w, double* out4,
                    cudaStream_t s);
int rbl_k_sort(rbl_ctx* c, const double* m, int64_t n, double* sorted_out, int32_t* perm_out, cudaStream_t s,
               const int32_t* prev_perm = nullptr);
int rbl_ss_buckets(int64_t n);
size_t rbl_ss_slots(int64_t n);
int rbl_sort_tiles(int64_t n);
int rbl_pav_chunk_log2();
int rbl_k_prefix(rbl_ctx* c, const double* x, int64_t n, double* loc_hi, double* loc_lo, double* tot_hi,
                 double* tot_lo, double* off_hi, double* off_lo, cudaStream_t s);
int rbl_k_segments(rbl_ctx* c, cudaStream_t s);
int rbl_pav_max_seg();
size_t rbl_pav_segblocks_bytes();
int rbl_k_pav(rbl_ctx* c, int loss, const double* m_sorted, double rho, double* z_sorted, cudaStream_t s);
int rbl_k_pass_multi(rbl_ctx* c, const double* D, const double* x, int64_t xstride, const double* b, double* r0,
                     double* r1, const FistaState* st, int ninst, double* red, int64_t red_stride, const double* c0part,
                     cudaStream_t s);
int rbl_k_fista_update_batch(rbl_ctx* c, int g0, int ng, cudaStream_t s);
int rbl_k_fista_result_batch(rbl_ctx* c, int B, double* w_out, double* r_out, cudaStream_t s);
size_t rbl_batch_smem(int64_t ld, int stages);
int rbl_batch_group();
int rbl_k_prox_elementwise(rbl_ctx* c, int loss, const double* sigma, const double* m, int64_t n, double rho,
                           double* out, cudaStream_t s);

int rbl_k_gram_fista_init(rbl_ctx* c, const double* G, const double* w0, const double* red0, cudaStream_t s);
int rbl_k_gram_fista_steps(rbl_ctx* c, const double* G, int nsteps, cudaStream_t s);
int rbl_k_gram_result(rbl_ctx* c, double* w_out, cudaStream_t s);
int rbl_k_gram_eval(rbl_ctx* c, const double* G, const double* w0, const double* red0, const double* w,
                    double* red_out, cudaStream_t s);
size_t rbl_gram_scratch_doubles(rbl_ctx* c, int64_t nrows);
int rbl_gram_persist_config(rbl_ctx* c);
int rbl_k_gram_fista_run(rbl_ctx* c, const double* G, const double* w0, const double* red0, double lam, int thr_f32,
                         float L0, double tol, int max_iter, double* w_out, double* w_prev_out, int with_support,
                         cudaStream_t s);
int rbl_k_lasso_cd_gram(rbl_ctx* c, const double* G, const double* w_ref, const double* red0, double l1, double tol,
                        int max_iter, double* w_out, double* info, cudaStream_t s);
int rbl_k_gram_build(rbl_ctx* c, const double* D, int64_t nrows, int accumulate, double* G, double* scratch,
                     cudaStream_t s);
int rbl_k_standardize_scratch_doubles(int num_sms, int64_t ld, int64_t* out);
int rbl_k_standardize(int num_sms, double* X, int64_t n, int64_t ld, double* mean, double* scale, double* scratch,
                      cudaStream_t s);
int rbl_k_gather_rows(int num_sms, const double* X, int64_t ld_in, const int64_t* idx, int64_t n_out, int64_t d,
                      double* out, int64_t ld_out, cudaStream_t s);
size_t rbl_k_metrics_scratch_bytes(int num_sms);
int rbl_k_test_metrics(int num_sms, const double* X, int64_t n, int64_t d, int64_t ld, const double* w,
                       const double* y, const int32_t* group, int loss, double threshold, double* out16,
                       void* scratch, cudaStream_t s);
int rbl_k_dual_sparse(rbl_ctx* c, const double* D, const double* Dt, const double* w, const double* z, double* Dw,
                      double* lam, double rho, int cap, int support_ready, cudaStream_t s);
int rbl_k_transpose(rbl_ctx* c, const double* D, double* Dt, cudaStream_t s);
int rbl_k_dual_finalize(rbl_ctx* c, int cap, const double* w, const double* w_prev, double* out8, double* w_copy,
                        cudaStream_t s);
int rbl_k_finalize(rbl_ctx* c, const double* part, int np, const double* w, const double* w_prev, double* out4,
                   cudaStream_t s);

static thread_local char g_err[512] = "";
long long g_rbl_launches = 0;
// measured on B200 inside the replayed iteration graph: 1409.8 it/s with programmatic dependent launch vs 1409.5
// without — the graph already removes the launch gaps it would hide, so it is opt-in (RBL_PDL=1)
int g_rbl_pdl = [] {
    const char* e = getenv("RBL_PDL");
    return (e && e[0] == '1') ? 1 : 0;
}();

void rbl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

// Handle scratch is ONE slab: ctx_alloc runs twice, first measuring (alloc_mode 1), then carving pointers out of a
// single cudaMalloc + cudaMemset (alloc_mode 2) — ~60 separate cudaMalloc/cudaMemset pairs used to cost ~0.1 s of
// every solver construction.  Later additions (batched buffers) allocate directly (alloc_mode 0) and are tracked.
template <class T>
int dev_alloc(rbl_ctx* c, T** p, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (c->alloc_mode == 1) {
        c->slab_bytes += bytes;
        *p = nullptr;
        return RBL_OK;
    }
    if (c->alloc_mode == 2) {
        *p = reinterpret_cast<T*>(c->slab + c->slab_off);
        c->slab_off += bytes;
        return RBL_OK;
    }
    RBL_CUDA(cudaMalloc((void**)p, bytes));
    RBL_CUDA(cudaMemset(*p, 0, bytes));
    if (c->n_extra < 32) c->extra[c->n_extra++] = (void*)*p;
    c->bytes += bytes;
    return RBL_OK;
}

#define RBL_TRY(x)                  \
    do {                            \
        int rc__ = (x);             \
        if (rc__ != RBL_OK) return rc__; \
    } while (0)

int ctx_alloc(rbl_ctx* c) {
    const size_t nl = (size_t)c->n_local, ng = (size_t)c->n_global, ld = (size_t)c->ld;
    RBL_TRY(dev_alloc(c, &c->gpart, (size_t)c->pass_grid * ld));
    RBL_TRY(dev_alloc(c, &c->sspart, (size_t)c->pass_grid));
    RBL_TRY(dev_alloc(c, &c->vpart, (size_t)4 * c->vec_grid));
    RBL_TRY(dev_alloc(c, &c->fista, 1));
    if (c->alloc_mode != 1) RBL_CUDA(cudaMallocHost((void**)&c->fista_host, 2 * sizeof(FistaState)));
    RBL_TRY(dev_alloc(c, &c->pow_tab, 128));
    RBL_TRY(dev_alloc(c, &c->beta, ld + 8));
    RBL_TRY(dev_alloc(c, &c->beta_p, ld + 8));
    RBL_TRY(dev_alloc(c, &c->beta_prev, ld + 8));
    RBL_TRY(dev_alloc(c, &c->g_p, ld + 8));
    RBL_TRY(dev_alloc(c, &c->g_prev, ld + 8));
    RBL_TRY(dev_alloc(c, &c->rbuf[0], nl));
    RBL_TRY(dev_alloc(c, &c->rbuf[1], nl));
    RBL_TRY(dev_alloc(c, &c->red_own, ld + 8));
    c->red = c->red_own;
    RBL_TRY(dev_alloc(c, &c->c0part, (size_t)c->vec_grid));
    RBL_TRY(dev_alloc(c, &c->gq_prev, ld + 8));
    RBL_TRY(dev_alloc(c, &c->gxs, 2 * ld + 8));
    RBL_TRY(dev_alloc(c, &c->gvu, 2 * ld + 8));
    RBL_TRY(dev_alloc(c, &c->gticket, 64));
    RBL_TRY(dev_alloc(c, &c->gvu2, 16 * ld + 8));
    RBL_TRY(dev_alloc(c, &c->act_cta_count, (size_t)c->vec_grid + 8));
    RBL_TRY(dev_alloc(c, &c->act_row, nl + 8));
    RBL_TRY(dev_alloc(c, &c->act_delta, nl + 8));
    RBL_TRY(dev_alloc(c, &c->act_total, 16));
    RBL_TRY(dev_alloc(c, &c->sup_idx, ld + 8));
    RBL_TRY(dev_alloc(c, &c->sup_val, ld + 8));
    RBL_TRY(dev_alloc(c, &c->sup_nnz, 64));
    c->sort_tiles = rbl_sort_tiles(c->n_global);
    RBL_TRY(dev_alloc(c, &c->keysA, ng));
    RBL_TRY(dev_alloc(c, &c->keysB, ng));
    RBL_TRY(dev_alloc(c, &c->valsA, ng));
    RBL_TRY(dev_alloc(c, &c->valsB, ng));
    RBL_TRY(dev_alloc(c, &c->tile_hist, (size_t)256 * c->sort_tiles + 256 + 8));
    RBL_TRY(dev_alloc(c, &c->sort_counts, (size_t)256 * c->num_sms));
    c->ss_nb = rbl_ss_buckets(c->n_global);
    RBL_TRY(dev_alloc(c, &c->ss_bkey, rbl_ss_slots(c->n_global)));
    RBL_TRY(dev_alloc(c, &c->ss_bval, rbl_ss_slots(c->n_global)));
    RBL_TRY(dev_alloc(c, &c->ss_count, (size_t)4096 + 8));
    RBL_TRY(dev_alloc(c, &c->ss_spl, (size_t)4096 + 8));
    RBL_TRY(dev_alloc(c, &c->ss_flag, 16));
    c->chunk_log2 = rbl_pav_chunk_log2();
    c->nchunks = (c->n_global + ((int64_t)1 << c->chunk_log2) - 1) >> c->chunk_log2;
    const size_t nch = (size_t)c->nchunks;
    RBL_TRY(dev_alloc(c, &c->ps_loc_hi, ng + 1));
    RBL_TRY(dev_alloc(c, &c->ps_loc_lo, ng + 1));
    RBL_TRY(dev_alloc(c, &c->ps_off_hi, nch + 1));
    RBL_TRY(dev_alloc(c, &c->ps_off_lo, nch + 1));
    RBL_TRY(dev_alloc(c, &c->ps_tot_hi, nch + 1));
    RBL_TRY(dev_alloc(c, &c->ps_tot_lo, nch + 1));
    RBL_TRY(dev_alloc(c, &c->pm_loc_hi, ng + 1));
    RBL_TRY(dev_alloc(c, &c->pm_loc_lo, ng + 1));
    RBL_TRY(dev_alloc(c, &c->pm_off_hi, nch + 1));
    RBL_TRY(dev_alloc(c, &c->pm_off_lo, nch + 1));
    RBL_TRY(dev_alloc(c, &c->ch_tot_hi, nch + 1));
    RBL_TRY(dev_alloc(c, &c->ch_tot_lo, nch + 1));
    RBL_TRY(dev_alloc(c, &c->node_cnt, nch + 64));
    RBL_TRY(dev_alloc(c, &c->sigma, ng));
    RBL_TRY(dev_alloc(c, &c->seg_count, 16));
    RBL_TRY(dev_alloc(c, &c->seg_bounds, (size_t)rbl_pav_max_seg() + 8));
    RBL_TRY(dev_alloc(c, (unsigned char**)&c->seg_blocks, rbl_pav_segblocks_bytes()));
    RBL_TRY(dev_alloc(c, &c->obj_tmp, ng));
    return RBL_OK;
}

void ctx_free(rbl_ctx* c) {
    if (c->slab) cudaFree(c->slab);
    for (int i = 0; i < c->n_extra; ++i)
        if (c->extra[i]) cudaFree(c->extra[i]);
    if (c->fista_host) cudaFreeHost(c->fista_host);
    if (c->eval_stage) cudaFreeHost(c->eval_stage);
    if (c->eval_w) cudaFree(c->eval_w);
    if (c->eval_red) cudaFree(c->eval_red);
    if (c->bfista_host) cudaFreeHost(c->bfista_host);
}

int ctx_alloc_slab(rbl_ctx* c) {
    c->alloc_mode = 1;
    c->slab_bytes = 0;
    int rc = ctx_alloc(c);
    if (rc != RBL_OK) return rc;
    RBL_CUDA(cudaMalloc((void**)&c->slab, c->slab_bytes));
    RBL_CUDA(cudaMemset(c->slab, 0, c->slab_bytes));
    c->bytes += c->slab_bytes;
    c->alloc_mode = 2;
    c->slab_off = 0;
    rc = ctx_alloc(c);
    c->alloc_mode = 0;
    return rc;
}

inline cudaStream_t S(rbl_stream_t s) { return (cudaStream_t)s; }

}  // namespace

#define RBL_ENTER(h)                                         \
    RBL_REQUIRE((h) != nullptr, "null handle");              \
    RBL_CUDA(cudaSetDevice((h)->device))

extern "C" {

int rbl_version(void) { return RBL_ABI_VERSION; }

const char* rbl_last_error(void) { return g_err; }

int64_t rbl_launch_count(void) { return (int64_t)g_rbl_launches; }

int rbl_create(rbl_handle_t* out, int device, int64_t n_local, int64_t n_global, int64_t row_lo, int32_t d,
               int64_t ld) {
    RBL_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    RBL_REQUIRE(n_local > 0 && n_global >= n_local && row_lo >= 0 && row_lo + n_local <= n_global,
                "bad row partition: n_local=%lld n_global=%lld row_lo=%lld", (long long)n_local, (long long)n_global,
                (long long)row_lo);
    RBL_REQUIRE(n_global < ((int64_t)1 << 31), "n_global must fit int32 permutation indices");
    RBL_REQUIRE(d > 0 && ld >= d, "bad d/ld: d=%d ld=%lld", d, (long long)ld);
    int ndev = 0;
    RBL_CUDA(cudaGetDeviceCount(&ndev));
    RBL_REQUIRE(device >= 0 && device < ndev, "no CUDA device %d (found %d); this library has no CPU path", device,
                ndev);
    RBL_CUDA(cudaSetDevice(device));
    // attributes one by one: cudaGetDeviceProperties fills ~100 fields and costs tens of milliseconds
    int cc_major = 0, cc_minor = 0, num_sms = 0;
    RBL_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    RBL_CUDA(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
    RBL_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    if (cc_major < 10) {
        rbl_set_error("device %d is sm_%d%d; librbl_b200 is built for sm_100a (B200) only", device, cc_major,
                      cc_minor);
        return RBL_ERR_UNSUPPORTED;
    }
    rbl_ctx* c = new rbl_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->num_sms = num_sms;
    c->n_local = n_local;
    c->n_global = n_global;
    c->row_lo = row_lo;
    c->d = d;
    c->ld = ld;
    c->vec_grid = c->num_sms * 4;
    int rc = rbl_pass_configure(c);
    if (rc == RBL_OK) rc = ctx_alloc_slab(c);
    if (rc != RBL_OK) {
        ctx_free(c);
        delete c;
        return rc;
    }
    *out = c;
    return RBL_OK;
}

int rbl_set_storage(rbl_handle_t h, int elem_bytes) {
    RBL_ENTER(h);
    RBL_REQUIRE(elem_bytes == 8 || elem_bytes == 4, "elem_bytes must be 8 (fp64) or 4 (fp32), got %d", elem_bytes);
    RBL_REQUIRE(h->batch_cap == 0 || elem_bytes == 8, "the batched multi-RHS pass reads fp64 storage only");
    const int old = h->esz;
    h->esz = elem_bytes;
    const int rc = rbl_pass_configure(h);  // row tiles are sized in bytes
    if (rc != RBL_OK) {
        h->esz = old;
        rbl_pass_configure(h);
    }
    return rc;
}

int rbl_set_pass_grid(rbl_handle_t h, int grid) {
    RBL_REQUIRE(h != nullptr, "null handle");
    RBL_REQUIRE(grid >= 1, "grid must be positive");
    h->pass_grid = grid > h->num_sms ? h->num_sms : grid;  // the partial buffers were sized for num_sms CTAs
    return RBL_OK;
}

int rbl_destroy(rbl_handle_t h) {
    if (!h) return RBL_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    ctx_free(h);
    delete h;
    return RBL_OK;
}

int rbl_info(rbl_handle_t h, int64_t* o) {
    RBL_REQUIRE(h && o, "null argument");
    o[0] = h->num_sms;
    o[1] = h->pass_grid;
    o[2] = h->pass_rows;
    o[3] = h->pass_stages;
    o[4] = (int64_t)h->pass_smem;
    o[5] = (int64_t)h->bytes;
    o[6] = h->vec_grid;
    o[7] = (int64_t)1 << h->chunk_log2;
    return RBL_OK;
}

int rbl_build_design(rbl_handle_t h, const double* X, int64_t ldx, const double* y, double* D, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(X && y && D && ldx >= h->d, "bad arguments");
    return rbl_k_build_design(h, X, ldx, y, D, h->n_local, S(stream));
}

int rbl_build_design_rows(rbl_handle_t h, const double* X, int64_t ldx, const double* y, double* D, int64_t nrows,
                          rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(X && y && D && ldx >= h->d && nrows >= 0 && nrows <= h->n_local, "bad arguments");
    return nrows ? rbl_k_build_design(h, X, ldx, y, D, nrows, S(stream)) : RBL_OK;
}

int rbl_set_spectrum(rbl_handle_t h, const double* sigma, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(sigma != nullptr, "sigma is null");
    RBL_CUDA(cudaMemcpyAsync(h->sigma, sigma, (size_t)h->n_global * sizeof(double), cudaMemcpyDeviceToDevice,
                             S(stream)));
    RBL_TRY(rbl_k_prefix(h, h->sigma, h->n_global, h->ps_loc_hi, h->ps_loc_lo, h->ps_tot_hi, h->ps_tot_lo,
                         h->ps_off_hi, h->ps_off_lo, S(stream)));
    RBL_TRY(rbl_k_segments(h, S(stream)));
    h->has_sigma = 1;
    return RBL_OK;
}

int rbl_matvec(rbl_handle_t h, const double* D, const double* x, double* out, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && x && out, "null argument");
    return rbl_launch_pass(h, RBL_PASS_MATVEC, D, x, nullptr, out, nullptr, nullptr, S(stream));
}

int rbl_margins(rbl_handle_t h, const double* Dw, const double* lam, double rho, double* m, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(Dw && lam && m, "null argument");
    return rbl_k_margins(h, Dw, lam, rho, m, S(stream));
}

int rbl_sort_margins(rbl_handle_t h, const double* m, double* m_sorted, int32_t* perm, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(m && (m_sorted || perm), "null argument");
    return rbl_k_sort(h, m, h->n_global, m_sorted, perm, S(stream));
}

int rbl_sort_margins_near(rbl_handle_t h, const double* m, const int32_t* prev_perm, double* m_sorted,
                          int32_t* perm, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(m && (m_sorted || perm), "null argument");
    return rbl_k_sort(h, m, h->n_global, m_sorted, perm, S(stream), prev_perm);
}

int rbl_pav_prox(rbl_handle_t h, int loss, const double* m_sorted, double rho, double* z_sorted,
                 rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(m_sorted && z_sorted, "null argument");
    RBL_REQUIRE(h->has_sigma, "rbl_set_spectrum must be called before rbl_pav_prox");
    RBL_REQUIRE(loss == RBL_LOSS_BINARY_CROSS_ENTROPY || loss == RBL_LOSS_HINGE, "unknown loss id %d", loss);
    RBL_REQUIRE(rho > 0.0, "rho must be positive");
    return rbl_k_pav(h, loss, m_sorted, rho, z_sorted, S(stream));
}

int rbl_bind_scalars(rbl_handle_t h, const double* d_scal) {
    RBL_REQUIRE(h != nullptr, "null handle");
    h->scal = d_scal;
    return RBL_OK;
}

int rbl_sort_config(rbl_handle_t h, int legacy) {
    RBL_REQUIRE(h != nullptr, "null handle");
    h->sort_legacy = (legacy & 1) ? 1 : 0;
    h->ss_off = (legacy & 2) ? 1 : 0;
    h->ss_row_order = (legacy & 4) ? 1 : 0;
    return RBL_OK;
}

int rbl_sort_stats(rbl_handle_t h, rbl_stream_t stream, int32_t* h_out) {
    RBL_ENTER(h);
    RBL_REQUIRE(h_out != nullptr, "null argument");
    int tmp[8];
    RBL_CUDA(cudaMemcpyAsync(tmp, h->ss_flag, sizeof(tmp), cudaMemcpyDeviceToHost, S(stream)));
    RBL_CUDA(cudaStreamSynchronize(S(stream)));
    h_out[0] = h->ss_nb;
    h_out[1] = tmp[4];  // route of the last hinted call: 1 buckets, 2 LSD fallback, 0 none yet
    h_out[2] = tmp[5];  // its largest bucket
    h_out[3] = tmp[0];  // overflow flag right now (0 between calls)
    return RBL_OK;
}

int rbl_sort_debug(rbl_handle_t h, uint64_t* d_stamps) {
    RBL_REQUIRE(h != nullptr, "null handle");
    h->sort_dbg = (unsigned long long*)d_stamps;
    return RBL_OK;
}

int rbl_pav_config(rbl_handle_t h, int force_tree, int32_t* h_nseg) {
    RBL_REQUIRE(h != nullptr, "null handle");
    h->force_tree = (force_tree & 1) ? 1 : 0;
    h->pav_no_hints = (force_tree & 2) ? 1 : 0;
    if (h_nseg) *h_nseg = h->nseg;
    return RBL_OK;
}

int rbl_prox_elementwise(rbl_handle_t h, int loss, const double* sigma, const double* m, int64_t n, double rho,
                         double* out, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(sigma && m && out && n >= 0, "bad arguments");
    RBL_REQUIRE(loss == RBL_LOSS_BINARY_CROSS_ENTROPY || loss == RBL_LOSS_HINGE, "unknown loss id %d", loss);
    RBL_REQUIRE(rho > 0.0, "rho must be positive");
    return rbl_k_prox_elementwise(h, loss, sigma, m, n, rho, out, S(stream));
}

int rbl_ehrm_candidate_sums(rbl_handle_t h, const double* m_sorted, const double* sigma_a, const double* sigma_b,
                            double B, double rho, double* out2, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(m_sorted && sigma_a && sigma_b && out2, "null argument");
    RBL_REQUIRE(rho > 0.0, "rho must be positive");
    return rbl_k_ehrm_sums(h, m_sorted, sigma_a, sigma_b, B, rho, out2, S(stream));
}

int rbl_scatter_z(rbl_handle_t h, const double* z_sorted, const int32_t* perm, int use_clip, double clip,
                  const double* lam, double rho, double* z, double* b, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(z_sorted && perm && z && (b == nullptr || lam != nullptr), "null argument");
    RBL_REQUIRE(use_clip >= 0 && use_clip <= 2, "use_clip must be 0 (none), 1 (max(B, .)) or 2 (min(B, .))");
    return rbl_k_scatter(h, z_sorted, perm, use_clip, clip, lam, rho, z, b, S(stream));
}

int rbl_scatter_active(rbl_handle_t h, const double* z_sorted, const double* m_sorted, const int32_t* perm,
                       int use_clip, double clip, const double* lam, double rho, double* z, double* b,
                       rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(z_sorted && m_sorted && perm && z && (b == nullptr || lam != nullptr), "null argument");
    RBL_REQUIRE(use_clip >= 0 && use_clip <= 2, "use_clip must be 0 (none), 1 (max(B, .)) or 2 (min(B, .))");
    return rbl_k_scatter_active(h, z_sorted, m_sorted, perm, use_clip, clip, lam, rho, z, b, S(stream));
}

int rbl_grad_pass(rbl_handle_t h, const double* D, const double* w0, const double* b, double* r, int64_t dense_above,
                  double* red, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && w0 && b && r && red, "null argument");
    const int cap = (int)(dense_above < 0 ? 0 : (dense_above > h->n_local ? h->n_local : dense_above));
    // exactly one of the two does the work, chosen on the device from the active-row count
    RBL_TRY(rbl_launch_gather(h, D, h->act_row, h->act_delta, h->act_total, cap, S(stream)));
    RBL_TRY(rbl_launch_pass(h, RBL_PASS_FUSED, D, w0, b, r, nullptr, nullptr, S(stream), nullptr, 0.0, h->act_total,
                            cap));
    RBL_TRY(rbl_k_reduce_partials(h, 0, nullptr, S(stream)));
    if (red != h->red)  // rbl_fista_bind_red(h, red) makes the reduction land in the caller's buffer directly
        RBL_CUDA(cudaMemcpyAsync(red, h->red, (size_t)(h->d + 1) * sizeof(double), cudaMemcpyDeviceToDevice,
                                 S(stream)));
    return RBL_OK;
}

int rbl_gather_only(rbl_handle_t h, const double* D, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D != nullptr, "null argument");
    return rbl_launch_gather(h, D, h->act_row, h->act_delta, h->act_total, (int)h->n_local, S(stream));
}

int rbl_active_count(rbl_handle_t h, int32_t* h_count, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(h_count != nullptr, "null argument");
    RBL_CUDA(cudaMemcpyAsync(h_count, h->act_total, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    RBL_CUDA(cudaStreamSynchronize(S(stream)));
    return RBL_OK;
}

int rbl_fused_pass(rbl_handle_t h, const double* D, const double* x, const double* b, double* r, double* red,
                   rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && x && b && r && red, "null argument");
    RBL_TRY(rbl_launch_pass(h, RBL_PASS_FUSED, D, x, b, r, nullptr, nullptr, S(stream)));
    RBL_TRY(rbl_k_reduce_partials(h, 0, nullptr, S(stream)));
    if (red != h->red)
        RBL_CUDA(cudaMemcpyAsync(red, h->red, (size_t)(h->d + 2) * sizeof(double), cudaMemcpyDeviceToDevice,
                                 S(stream)));
    return RBL_OK;
}

int rbl_fista_config(rbl_handle_t h, const float* h_pow_tab) {
    RBL_ENTER(h);
    RBL_REQUIRE(h_pow_tab != nullptr, "null table");
    RBL_CUDA(cudaMemcpy(h->pow_tab, h_pow_tab, 128 * sizeof(float), cudaMemcpyHostToDevice));
    return RBL_OK;
}

int rbl_fista_begin(rbl_handle_t h, const double* w0, double lam, int thr_f32, float L0, double tol, int max_iter,
                    rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(w0 != nullptr && max_iter > 0, "bad arguments");
    FistaState* st = &h->fista_host[1];  // staging slot ([0] is the poll mirror)
    memset(st, 0, sizeof(*st));
    st->t = 1.0;
    st->tol = tol;
    st->lam = lam;
    st->L_prev = L0;
    st->L_cur = L0;
    st->k = -1;
    st->max_iter = max_iter;
    st->thr_f32 = thr_f32;
    RBL_CUDA(cudaMemcpyAsync(h->fista, st, sizeof(FistaState), cudaMemcpyHostToDevice, S(stream)));
    RBL_CUDA(cudaMemcpyAsync(h->beta, w0, (size_t)h->d * sizeof(double), cudaMemcpyDeviceToDevice, S(stream)));
    return RBL_OK;
}

int rbl_fista_pass(rbl_handle_t h, const double* D, const double* b, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && b, "null argument");
    RBL_TRY(rbl_launch_pass(h, RBL_PASS_FISTA, D, h->beta, b, nullptr, h->fista, h->rbuf, S(stream)));
    return rbl_k_reduce_partials(h, 1, h->fista, S(stream));
}

int rbl_fista_bind_red(rbl_handle_t h, double* red) {
    RBL_REQUIRE(h != nullptr, "null handle");
    h->red = red ? red : h->red_own;
    return RBL_OK;
}

int rbl_fista_update(rbl_handle_t h, rbl_stream_t stream) {
    RBL_ENTER(h);
    return rbl_k_fista_update(h, S(stream));
}

int rbl_fista_steps(rbl_handle_t h, const double* D, const double* b, int nsteps, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && b && nsteps > 0, "bad arguments");
    for (int i = 0; i < nsteps; ++i) {
        RBL_TRY(rbl_launch_pass(h, RBL_PASS_FISTA, D, h->beta, b, nullptr, h->fista, h->rbuf, S(stream)));
        RBL_TRY(rbl_k_reduce_partials(h, 1, h->fista, S(stream)));
        RBL_TRY(rbl_k_fista_update(h, S(stream)));
    }
    return RBL_OK;
}

int rbl_fista_poll(rbl_handle_t h, rbl_stream_t stream, int32_t* hi, double* hd) {
    RBL_ENTER(h);
    RBL_REQUIRE(hi && hd, "null argument");
    RBL_CUDA(cudaMemcpyAsync(&h->fista_host[0], h->fista, sizeof(FistaState), cudaMemcpyDeviceToHost, S(stream)));
    RBL_CUDA(cudaStreamSynchronize(S(stream)));
    const FistaState& st = h->fista_host[0];
    hi[0] = st.done;
    hi[1] = st.k;
    hi[2] = st.passes;
    hi[3] = st.trials;
    hi[4] = st.i_k;
    hi[5] = st.cur;
    hd[0] = st.crit;
    hd[1] = (double)st.L_prev;
    hd[2] = st.t;
    hd[3] = st.ss_last;
    return RBL_OK;
}

int rbl_fista_result(rbl_handle_t h, double* w_out, double* r_out, rbl_stream_t stream) {
    RBL_ENTER(h);
    return rbl_k_fista_result(h, w_out, r_out, S(stream));
}

// ---- batched mode (K10) ---------------------------------------------------------------------------
int rbl_batch_create(rbl_handle_t h, int B) {
    RBL_ENTER(h);
    RBL_REQUIRE(B > 0 && B <= 4096, "bad batch size %d", B);
    RBL_REQUIRE(h->esz != 4, "the batched multi-RHS pass reads fp64 storage only (rbl_set_storage)");
    RBL_REQUIRE(h->batch_cap == 0, "batched buffers already created for this handle");
    RBL_REQUIRE(h->ld <= 1024, "batched mode supports d <= 1024 (got %d)", h->d);
    int dev_max = 0;
    RBL_CUDA(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    int stages = 3;
    while (stages > 1 && rbl_batch_smem(h->ld, stages) > (size_t)dev_max) --stages;
    RBL_REQUIRE(rbl_batch_smem(h->ld, stages) <= (size_t)dev_max, "row tiles do not fit in shared memory");
    h->batch_stages = stages;
    const size_t vs = (size_t)h->ld + 8, nl = (size_t)h->n_local;
    RBL_TRY(dev_alloc(h, &h->bfista, (size_t)B));
    RBL_CUDA(cudaMallocHost((void**)&h->bfista_host, 2 * (size_t)B * sizeof(FistaState)));
    RBL_TRY(dev_alloc(h, &h->bbeta, B * vs));
    RBL_TRY(dev_alloc(h, &h->bbeta_p, B * vs));
    RBL_TRY(dev_alloc(h, &h->bbeta_prev, B * vs));
    RBL_TRY(dev_alloc(h, &h->bg_p, B * vs));
    RBL_TRY(dev_alloc(h, &h->bg_prev, B * vs));
    RBL_TRY(dev_alloc(h, &h->brbuf[0], B * nl));
    RBL_TRY(dev_alloc(h, &h->brbuf[1], B * nl));
    RBL_TRY(dev_alloc(h, &h->bred, B * vs));
    RBL_TRY(dev_alloc(h, &h->bgpart, (size_t)h->pass_grid * rbl_batch_group() * h->ld));
    RBL_TRY(dev_alloc(h, &h->bsspart, (size_t)h->pass_grid * rbl_batch_group()));
    RBL_TRY(dev_alloc(h, &h->bc0part, (size_t)B * h->vec_grid));
    h->batch_cap = B;
    return RBL_OK;
}

int rbl_fista_batch_begin(rbl_handle_t h, int B, const double* w0s, const double* h_lams, const int32_t* h_thr_f32,
                          float L0, double tol, int max_iter, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(B > 0 && B <= h->batch_cap, "batch of %d exceeds rbl_batch_create capacity %d", B, h->batch_cap);
    RBL_REQUIRE(w0s && h_lams && h_thr_f32 && max_iter > 0, "bad arguments");
    FistaState* st = h->bfista_host + h->batch_cap;  // staging half
    for (int g = 0; g < B; ++g) {
        memset(&st[g], 0, sizeof(FistaState));
        st[g].t = 1.0;
        st[g].tol = tol;
        st[g].lam = h_lams[g];
        st[g].L_prev = L0;
        st[g].L_cur = L0;
        st[g].k = -1;
        st[g].max_iter = max_iter;
        st[g].thr_f32 = h_thr_f32[g];
    }
    RBL_CUDA(cudaMemcpyAsync(h->bfista, st, (size_t)B * sizeof(FistaState), cudaMemcpyHostToDevice, S(stream)));
    RBL_CUDA(cudaMemcpy2DAsync(h->bbeta, ((size_t)h->ld + 8) * sizeof(double), w0s, (size_t)h->d * sizeof(double),
                               (size_t)h->d * sizeof(double), (size_t)B, cudaMemcpyDeviceToDevice, S(stream)));
    return RBL_OK;
}

int rbl_fista_batch_steps(rbl_handle_t h, int B, const double* D, const double* bs, int nsteps,
                          rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(B > 0 && B <= h->batch_cap && D && bs && nsteps > 0, "bad arguments");
    const int G = rbl_batch_group();
    const size_t vs = (size_t)h->ld + 8, nl = (size_t)h->n_local;
    for (int i = 0; i < nsteps; ++i) {
        for (int g0 = 0; g0 < B; g0 += G) {
            const int ng = (B - g0 < G) ? (B - g0) : G;
            RBL_TRY(rbl_k_pass_multi(h, D, h->bbeta + g0 * vs, (int64_t)vs, bs + g0 * nl, h->brbuf[0] + g0 * nl,
                                     h->brbuf[1] + g0 * nl, h->bfista + g0, ng, h->bred + g0 * vs, (int64_t)vs,
                                     h->bc0part + (size_t)g0 * h->vec_grid, S(stream)));
            RBL_TRY(rbl_k_fista_update_batch(h, g0, ng, S(stream)));
        }
    }
    return RBL_OK;
}

int rbl_fista_batch_poll(rbl_handle_t h, int B, rbl_stream_t stream, int32_t* h_done, int32_t* h_iters,
                         int32_t* h_passes, double* h_L) {
    RBL_ENTER(h);
    RBL_REQUIRE(B > 0 && B <= h->batch_cap && h_done && h_iters && h_passes, "bad arguments");
    RBL_CUDA(cudaMemcpyAsync(h->bfista_host, h->bfista, (size_t)B * sizeof(FistaState), cudaMemcpyDeviceToHost,
                             S(stream)));
    RBL_CUDA(cudaStreamSynchronize(S(stream)));
    for (int g = 0; g < B; ++g) {
        h_done[g] = h->bfista_host[g].done;
        h_iters[g] = h->bfista_host[g].k;
        h_passes[g] = h->bfista_host[g].passes;
        if (h_L) h_L[g] = (double)h->bfista_host[g].L_prev;
    }
    return RBL_OK;
}

int rbl_fista_batch_result(rbl_handle_t h, int B, double* w_out, double* r_out, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(B > 0 && B <= h->batch_cap, "bad arguments");
    return rbl_k_fista_result_batch(h, B, w_out, r_out, S(stream));
}

// ---- Gram mode (gram_kernels.cu) ------------------------------------------------------------------
static int gram_rows(rbl_handle_t h, const double* D, int64_t nrows, int accumulate, double* G, cudaStream_t s) {
    double* scratch = nullptr;
    const size_t bytes = rbl_gram_scratch_doubles(h, nrows) * sizeof(double);
    RBL_CUDA(cudaMallocAsync((void**)&scratch, bytes, s));
    int rc = rbl_k_gram_build(h, D, nrows, accumulate, G, scratch, s);
    RBL_CUDA(cudaFreeAsync(scratch, s));
    return rc;
}

int rbl_gram_build(rbl_handle_t h, const double* D, double* G, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && G, "null argument");
    return gram_rows(h, D, h->n_local, 0, G, S(stream));
}

int rbl_gram_accumulate(rbl_handle_t h, const double* D_rows, int64_t nrows, int accumulate, double* G,
                        rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D_rows && G && nrows > 0 && nrows <= h->n_local, "bad arguments");
    return gram_rows(h, D_rows, nrows, accumulate, G, S(stream));
}

int rbl_gram_fista_begin(rbl_handle_t h, const double* G, const double* w0, const double* red0, double lam,
                         int thr_f32, float L0, double tol, int max_iter, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && w0 && red0 && max_iter > 0, "bad arguments");
    FistaState* st = &h->fista_host[1];
    memset(st, 0, sizeof(*st));
    st->t = 1.0;
    st->tol = tol;
    st->lam = lam;
    st->L_prev = L0;
    st->L_cur = L0;
    st->k = -1;
    st->max_iter = max_iter;
    st->thr_f32 = thr_f32;
    RBL_CUDA(cudaMemcpyAsync(h->fista, st, sizeof(FistaState), cudaMemcpyHostToDevice, S(stream)));
    return rbl_k_gram_fista_init(h, G, w0, red0, S(stream));
}

int rbl_gram_fista_run(rbl_handle_t h, const double* G, const double* w0, const double* red0, double lam, int thr_f32,
                       float L0, double tol, int max_iter, double* w_out, double* w_prev_out, int with_support,
                       rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && w0 && red0 && max_iter > 0, "bad arguments");
    RBL_REQUIRE(w_prev_out == nullptr || w_prev_out != w_out, "w_prev_out and w_out must not alias");
    if (!rbl_gram_persist_config(h)) {
        rbl_set_error("persistent FISTA kernel unavailable for d = %d (state does not fit in shared memory or no "
                      "cooperative launch); use rbl_gram_fista_begin/steps", h->d);
        return RBL_ERR_UNSUPPORTED;
    }
    return rbl_k_gram_fista_run(h, G, w0, red0, lam, thr_f32, L0, tol, max_iter, w_out, w_prev_out, with_support,
                                S(stream));
}

int rbl_gram_fista_persistent_ok(rbl_handle_t h) {
    if (!h) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return 0;
    return rbl_gram_persist_config(h);
}

int rbl_gram_fista_steps(rbl_handle_t h, const double* G, int nsteps, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && nsteps > 0 && h->gram_w0, "bad arguments (rbl_gram_fista_begin first)");
    return rbl_k_gram_fista_steps(h, G, nsteps, S(stream));
}

int rbl_gram_fista_result(rbl_handle_t h, double* w_out, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(w_out != nullptr, "null argument");
    return rbl_k_gram_result(h, w_out, S(stream));
}

int rbl_gram_eval(rbl_handle_t h, const double* G, const double* w0, const double* red0, const double* w,
                  double* red_out, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && w0 && red0 && w && red_out, "null argument");
    return rbl_k_gram_eval(h, G, w0, red0, w, red_out, S(stream));
}

int rbl_lasso_cd_gram(rbl_handle_t h, const double* G, const double* w_ref, const double* red0, double l1, double tol,
                      int max_iter, double* w_out, double* info3, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && w_ref && red0 && w_out, "null argument");
    RBL_REQUIRE(h->d <= 64, "rbl_lasso_cd_gram serves the reference's small-problem branch (d <= 60); d = %d", h->d);
    RBL_REQUIRE(w_out != w_ref, "w_out and w_ref must not alias");
    RBL_REQUIRE(l1 >= 0.0 && tol > 0.0 && max_iter > 0, "bad arguments");
    return rbl_k_lasso_cd_gram(h, G, w_ref, red0, l1, tol, max_iter, w_out, info3, S(stream));
}

int rbl_gram_eval_host(rbl_handle_t h, const double* G, const double* w0, const double* red0, const double* h_w,
                       double* h_red_out, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && w0 && red0 && h_w && h_red_out, "null argument");
    const size_t nw = (size_t)h->d, nr = (size_t)h->d + 2;
    if (!h->eval_stage) {  // pinned staging [w (ld + 8) | red (ld + 8)], device copy of w: created on first use
        RBL_CUDA(cudaMallocHost((void**)&h->eval_stage, 2 * ((size_t)h->ld + 8) * sizeof(double)));
        RBL_CUDA(cudaMalloc((void**)&h->eval_w, ((size_t)h->ld + 8) * sizeof(double)));
        RBL_CUDA(cudaMalloc((void**)&h->eval_red, ((size_t)h->ld + 8) * sizeof(double)));
    }
    double* st_w = h->eval_stage;
    double* st_r = h->eval_stage + h->ld + 8;
    memcpy(st_w, h_w, nw * sizeof(double));
    RBL_CUDA(cudaMemcpyAsync(h->eval_w, st_w, nw * sizeof(double), cudaMemcpyHostToDevice, S(stream)));
    RBL_TRY(rbl_k_gram_eval(h, G, w0, red0, h->eval_w, h->eval_red, S(stream)));
    RBL_CUDA(cudaMemcpyAsync(st_r, h->eval_red, nr * sizeof(double), cudaMemcpyDeviceToHost, S(stream)));
    RBL_CUDA(cudaStreamSynchronize(S(stream)));
    memcpy(h_red_out, st_r, nr * sizeof(double));
    return RBL_OK;
}

// ---- the smooth w-steps in the library (w_LBFGS.py:48-62): L-BFGS-B on the host (csrc/lbfgs_core.h) over device
// f/g evaluations on G.  f(w) = rho/2 ||D w - b||^2 + R(w), grad = rho D^T (D w - b) + R'(w) with
//   reg_kind 0: R = reg/2 ||w||^2                                        (wl2_fun / wl2_fun_gradient, :31-45)
//   reg_kind 1: R = reg/2 sum_j (w_j^2 / (2t) if |w_j| <= t else |w_j| - t/2)   (wl1_fun_smooth, :11-28)
int rbl_lbfgs_gram(rbl_handle_t h, const double* G, const double* w0, const double* red0, double rho, double reg,
                   int reg_kind, double t, int maxiter, double* h_w, double* d_w_out, int32_t* h_info,
                   rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(G && w0 && red0 && h_w && h_info, "null argument");
    RBL_REQUIRE(reg_kind == 0 || (reg_kind == 1 && t > 0.0), "reg_kind must be 0 (l2) or 1 (smoothed l1, t > 0)");
    RBL_REQUIRE(rho > 0.0 && maxiter > 0, "bad arguments");
    const int d = h->d;
    if (!h->eval_stage) {
        RBL_CUDA(cudaMallocHost((void**)&h->eval_stage, 2 * ((size_t)h->ld + 8) * sizeof(double)));
        RBL_CUDA(cudaMalloc((void**)&h->eval_w, ((size_t)h->ld + 8) * sizeof(double)));
        RBL_CUDA(cudaMalloc((void**)&h->eval_red, ((size_t)h->ld + 8) * sizeof(double)));
    }
    double* st_w = h->eval_stage;
    double* st_r = h->eval_stage + h->ld + 8;
    cudaStream_t s = S(stream);
    int cuda_rc = RBL_OK;
    auto fg = [&](const double* x, double* f, double* g) -> int {
        memcpy(st_w, x, (size_t)d * sizeof(double));
        if (cudaMemcpyAsync(h->eval_w, st_w, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
            return cuda_rc = RBL_ERR_CUDA;
        if ((cuda_rc = rbl_k_gram_eval(h, G, w0, red0, h->eval_w, h->eval_red, s)) != RBL_OK) return cuda_rc;
        if (cudaMemcpyAsync(st_r, h->eval_red, ((size_t)d + 2) * sizeof(double), cudaMemcpyDeviceToHost, s) !=
                cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess)
            return cuda_rc = RBL_ERR_CUDA;
        double R = 0.0;
        if (reg_kind == 0) {
            double ww = 0.0;
            for (int j = 0; j < d; ++j) ww += x[j] * x[j];
            R = 0.5 * reg * ww;
            for (int j = 0; j < d; ++j) g[j] = -rho * st_r[j] + reg * x[j];
        } else {
            double sq = 0.0, ab = 0.0;
            for (int j = 0; j < d; ++j) {
                const double a = fabs(x[j]);
                if (a <= t) {
                    sq += x[j] * x[j];
                    g[j] = -rho * st_r[j] + 0.5 * reg * x[j] / t;
                } else {
                    ab += a - 0.5 * t;
                    g[j] = -rho * st_r[j] + 0.5 * reg * (x[j] > 0.0 ? 1.0 : -1.0);
                }
            }
            R = 0.5 * 0.5 * reg * sq / t + 0.5 * reg * ab;
        }
        *f = 0.5 * rho * st_r[d] + R;
        return 0;
    };
    const rbl_lbfgs::Result res = rbl_lbfgs::minimize(d, h_w, fg, 10, maxiter);
    if (cuda_rc != RBL_OK) {
        rbl_set_error("rbl_lbfgs_gram: a device evaluation failed (%s)", cudaGetErrorString(cudaGetLastError()));
        return cuda_rc;
    }
    h_info[0] = res.nit;
    h_info[1] = res.nfev;
    h_info[2] = res.status;
    if (d_w_out) {
        memcpy(st_w, h_w, (size_t)d * sizeof(double));
        RBL_CUDA(cudaMemcpyAsync(d_w_out, st_w, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    return RBL_OK;
}

int rbl_build_transpose(rbl_handle_t h, const double* D, double* Dt, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && Dt, "null argument");
    return rbl_k_transpose(h, D, Dt, S(stream));
}

int rbl_dual_pass(rbl_handle_t h, const double* D, const double* Dt, const double* w, const double* w_prev,
                  const double* z, double* Dw, double* lam, double rho, int sparse_cap, int support_ready,
                  double* out8, double* w_copy, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(D && w && z && Dw && lam && out8, "null argument");
    const int cap = sparse_cap < 0 ? 0 : sparse_cap;
    // exactly one of the two kernels does the work, chosen on the device from nnz(w): no host round trip
    RBL_TRY(rbl_k_dual_sparse(h, D, Dt, w, z, Dw, lam, rho, cap, support_ready, S(stream)));
    RBL_TRY(rbl_launch_pass(h, RBL_PASS_DUAL, D, w, z, Dw, nullptr, nullptr, S(stream), lam, rho, h->sup_nnz, cap));
    return rbl_k_dual_finalize(h, cap, w, w_prev, out8, w_copy, S(stream));
}

// ---- native outer loop over a captured iteration graph (algorithms.py:119-157 host logic) -------------------
int rbl_admm_run(rbl_handle_t h, void* graph_exec, rbl_stream_t stream, double* h_scal, const double* h_out,
                 int32_t max_iters, double tol, double reg, int64_t num_row, int32_t num_feature,
                 int64_t dense_above, double rho, int32_t rho_is_pyfloat, rbl_run_stats* out) {
    RBL_ENTER(h);
    RBL_REQUIRE(graph_exec && h_scal && h_out && out && max_iters >= 0, "bad arguments");
    memset(out, 0, sizeof(*out));
    const double rho_cap = 217.0 * (double)num_feature;  // algorithms.py:153-157
    int pyfloat = rho_is_pyfloat ? 1 : 0;
    for (int it = 0; it < max_iters; ++it) {
        // lam = alpha * n with alpha = reg / (2 rho n), in the reference's operation order (:192-193,200)
        const double alpha = reg / (2.0 * rho * (double)num_row);
        h_scal[0] = rho;
        h_scal[1] = alpha * (double)num_row;
        h_scal[2] = pyfloat ? 1.0 : 0.0;  // python-float lam: float32 threshold quotient (NEP 50), else float64
        RBL_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, S(stream)));
        RBL_CUDA(cudaStreamSynchronize(S(stream)));
        const double primal = sqrt(h_out[0]), dual = sqrt(h_out[1]);
        out->iters = it + 1;
        out->primal = primal;
        out->dual = dual;
        out->nnz_last = (int32_t)h_out[4];
        if (h_out[5] != 0.0) ++out->sparse_dual; else ++out->dense_dual;
        out->fista_iters += (int64_t)h_out[6];
        out->fista_sweeps += (int64_t)h_out[7];
        out->last_sweeps = (int32_t)h_out[7];
        const int64_t rows = (int64_t)h_out[8];
        if (rows <= dense_above) {
            ++out->gathered;
            out->rows_read += rows;
        } else {
            out->rows_read += h->n_local;
        }
        if (primal < tol && dual < tol) {  // :137 — before the rho update, like the reference
            out->converged = 1;
            break;
        }
        rho = fmin(rho * (primal > 1e-2 ? 1.02 : 1.07), rho_cap);  // :153-157
        pyfloat = 0;                                               // np.min returns np.float64
    }
    out->rho = rho;
    out->rho_is_pyfloat = pyfloat;
    return RBL_OK;
}

// ---- native outer loop for the smooth (l2) problems: [graph: z-step + warm-start gradient pass] -> L-BFGS-B in the
// library -> [graph: dual pass + read-back], with the reference's stop test and rho schedule in between
int rbl_admm_run_l2(rbl_handle_t h, void* graph_pre, void* graph_dual, rbl_stream_t stream, double* h_scal,
                    const double* h_out, const double* G, const double* w0, const double* red0, double* h_w,
                    double* d_w, double reg, int32_t lbfgs_maxiter, int32_t max_iters, double tol, int32_t num_feature,
                    int64_t dense_above, double rho, rbl_run_stats* out) {
    RBL_ENTER(h);
    RBL_REQUIRE(graph_pre && graph_dual && h_scal && h_out && G && w0 && red0 && h_w && d_w && out && max_iters >= 0,
                "bad arguments");
    memset(out, 0, sizeof(*out));
    const double rho_cap = 217.0 * (double)num_feature;  // algorithms.py:153-157
    int32_t info[4];
    for (int it = 0; it < max_iters; ++it) {
        h_scal[0] = rho;
        RBL_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_pre, S(stream)));
        RBL_CUDA(cudaStreamSynchronize(S(stream)));  // h_w (pinned) now holds the warm start, red0 the gradient pass
        RBL_TRY(rbl_lbfgs_gram(h, G, w0, red0, rho, reg, 0, 0.0, lbfgs_maxiter, h_w, d_w, info, stream));
        RBL_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_dual, S(stream)));
        RBL_CUDA(cudaStreamSynchronize(S(stream)));
        const double primal = sqrt(h_out[0]), dual = sqrt(h_out[1]);
        out->iters = it + 1;
        out->primal = primal;
        out->dual = dual;
        out->nnz_last = (int32_t)h_out[4];
        if (h_out[5] != 0.0) ++out->sparse_dual; else ++out->dense_dual;
        out->fista_iters += info[0];   // L-BFGS iterations
        out->fista_sweeps += info[1];  // f/g evaluations (one sweep over G each)
        out->last_sweeps = info[1];
        const int64_t rows = (int64_t)h_out[8];
        if (rows <= dense_above) {
            ++out->gathered;
            out->rows_read += rows;
        } else {
            out->rows_read += h->n_local;
        }
        if (primal < tol && dual < tol) {
            out->converged = 1;
            break;
        }
        rho = fmin(rho * (primal > 1e-2 ? 1.02 : 1.07), rho_cap);
    }
    out->rho = rho;
    out->rho_is_pyfloat = 0;
    return RBL_OK;
}

int rbl_dual_update(rbl_handle_t h, const double* z, double* Dw, const double* b, const double* r,
                    int from_residual, double* lam, double rho, const double* w, const double* w_prev, double* out4,
                    rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(z && Dw && lam && out4, "null argument");
    RBL_REQUIRE(!from_residual || (b && r), "from_residual needs b and r");
    return rbl_k_dual(h, z, Dw, b, r, from_residual, lam, rho, w, w_prev, out4, S(stream));
}

int rbl_objective(rbl_handle_t h, int loss, const double* margins, const double* sigma, const double* w,
                  double* out4, rbl_stream_t stream) {
    RBL_ENTER(h);
    RBL_REQUIRE(margins && sigma && out4, "null argument");
    RBL_TRY(rbl_k_sort(h, margins, h->n_global, h->obj_tmp, nullptr, S(stream)));
    return rbl_k_objective(h, h->obj_tmp, sigma, loss, w, out4, S(stream));
}

// ---- upload from PAGEABLE host memory (a user's numpy array) -----------------------------------------------------
// cudaMemcpyAsync from pageable memory is staged by the driver through one small pinned buffer on the calling
// thread (~10 GB/s).  Here `nthreads` host threads copy 8 MB chunks into their own pinned slots (2 per thread)
// and each issues its DMA on its own stream, so the host-side memcpy runs at the memory system's multi-threaded
// bandwidth and overlaps the transfers; `stream` then waits for every chunk.  The pinned slots are created once
// per process and device.  Returns when all chunks are staged (the transfers may still be in flight on `stream`).
namespace {
constexpr size_t kUpChunk = (size_t)8 << 20;
constexpr int kUpMaxThreads = 16;
struct UploadLane {
    void* pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaStream_t s = nullptr;
};
struct UploadPool {
    int device = -1;
    UploadLane lane[kUpMaxThreads];
    int nready = 0;
};
UploadPool g_up[RBL_MAX_DEVICES];

int upload_prepare(UploadPool& up, int device, int nthreads) {
    up.device = device;
    for (int t = up.nready; t < nthreads; ++t) {
        UploadLane& L = up.lane[t];
        for (int k = 0; k < 2; ++k) {
            RBL_CUDA(cudaMallocHost(&L.pin[k], kUpChunk));
            RBL_CUDA(cudaEventCreateWithFlags(&L.done[k], cudaEventDisableTiming));
        }
        RBL_CUDA(cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking));
        up.nready = t + 1;
    }
    return RBL_OK;
}
}  // namespace

int rbl_h2d_pageable(int device, void* d_dst, const void* h_src, int64_t bytes, int nthreads, rbl_stream_t stream) {
    RBL_REQUIRE(d_dst && h_src && bytes >= 0, "bad arguments");
    RBL_REQUIRE(device >= 0 && device < RBL_MAX_DEVICES, "bad device %d", device);
    if (bytes == 0) return RBL_OK;
    RBL_CUDA(cudaSetDevice(device));
    if (nthreads < 1) nthreads = 1;
    if (nthreads > kUpMaxThreads) nthreads = kUpMaxThreads;
    const int64_t nchunks = (bytes + (int64_t)kUpChunk - 1) / (int64_t)kUpChunk;
    if (nchunks < nthreads) nthreads = (int)nchunks;
    UploadPool& up = g_up[device];
    RBL_TRY(upload_prepare(up, device, nthreads));
    // the destination may still be read by earlier work on `stream`: the lanes start after it
    cudaEvent_t start;
    RBL_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    RBL_CUDA(cudaEventRecord(start, S(stream)));
    std::vector<int> rc(nthreads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            if (cudaSetDevice(device) != cudaSuccess) { rc[t] = 1; return; }
            UploadLane& L = up.lane[t];
            if (cudaStreamWaitEvent(L.s, start, 0) != cudaSuccess) { rc[t] = 1; return; }
            int slot = 0;
            for (int64_t k = t; k < nchunks; k += nthreads, slot ^= 1) {
                const size_t off = (size_t)k * kUpChunk;
                const size_t len = (size_t)bytes - off < kUpChunk ? (size_t)bytes - off : kUpChunk;
                if (cudaEventSynchronize(L.done[slot]) != cudaSuccess) { rc[t] = 1; return; }  // slot free again
                memcpy(L.pin[slot], (const char*)h_src + off, len);
                if (cudaMemcpyAsync((char*)d_dst + off, L.pin[slot], len, cudaMemcpyHostToDevice, L.s) != cudaSuccess ||
                    cudaEventRecord(L.done[slot], L.s) != cudaSuccess) { rc[t] = 1; return; }
            }
        });
    }
    for (auto& x : th) x.join();
    RBL_CUDA(cudaEventDestroy(start));
    for (int t = 0; t < nthreads; ++t) {
        RBL_REQUIRE(rc[t] == 0, "upload lane %d failed: %s", t, cudaGetErrorString(cudaGetLastError()));
        for (int k = 0; k < 2; ++k) RBL_CUDA(cudaStreamWaitEvent(S(stream), up.lane[t].done[k], 0));
    }
    return RBL_OK;
}

// ---- CPT spectra of EHRM on the host (objective.py:148-164) ------------------------------------------------------
// a_i = distort((i+1)/n, 0.69) - distort(i/n, 0.69), b_i = distort((n-i)/n, 0.61) - distort((n-i-1)/n, 0.61) with
// distort(p, g) = p^g / (p^g + (1-p)^g)^(1/g).  The differences cancel ~log10(n) digits, so the LAST bit of every
// pow matters: the reference evaluates them one Python float at a time (libm pow); the same scalar libm calls in
// the same order are made here (a vectorised pow differs in the last ulp: 3e-9 relative in sigma at n = 4M).
static double cpt_distort(double p, double g) { return pow(p, g) / pow(pow(p, g) + pow(1.0 - p, g), 1.0 / g); }

int rbl_cpt_weights(int64_t n, int which, double* h_out) {
    RBL_REQUIRE(n > 0 && h_out != nullptr && (which == 0 || which == 1), "bad arguments");
    const double dn = (double)n;
    if (which == 0)
        for (int64_t i = 0; i < n; ++i)
            h_out[i] = cpt_distort((double)(i + 1) / dn, 0.69) - cpt_distort((double)i / dn, 0.69);
    else
        for (int64_t i = 0; i < n; ++i)
            h_out[i] = cpt_distort((double)(n - i) / dn, 0.61) - cpt_distort((double)(n - i - 1) / dn, 0.61);
    return RBL_OK;
}

// ---- test-set metrics (no handle: a test set has its own row count) -------------------------------------------
static int metrics_device(int device, int* num_sms) {
    int ndev = 0, cc_major = 0;
    RBL_CUDA(cudaGetDeviceCount(&ndev));
    RBL_REQUIRE(device >= 0 && device < ndev, "no CUDA device %d (found %d); this library has no CPU path", device,
                ndev);
    RBL_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    RBL_REQUIRE(cc_major >= 10, "device %d is not sm_100 class; librbl_b200 is built for sm_100a (B200) only", device);
    RBL_CUDA(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, device));
    return RBL_OK;
}

int rbl_metrics_scratch_bytes(int device, int64_t* bytes) {
    RBL_REQUIRE(bytes != nullptr, "null argument");
    int num_sms = 0;
    RBL_TRY(metrics_device(device, &num_sms));
    *bytes = (int64_t)rbl_k_metrics_scratch_bytes(num_sms);
    return RBL_OK;
}

int rbl_test_metrics(int device, const double* X, int64_t n, int64_t d, int64_t ld, const double* w,
                     const double* y, const int32_t* group, int loss, double threshold, double* out16,
                     void* scratch, rbl_stream_t stream) {
    RBL_REQUIRE(X && w && y && out16 && scratch, "null argument");
    RBL_REQUIRE(n > 0 && d > 0 && ld >= d, "bad shape: n=%lld d=%lld ld=%lld", (long long)n, (long long)d,
                (long long)ld);
    RBL_REQUIRE(loss == RBL_LOSS_BCE || loss == RBL_LOSS_HINGE, "unknown loss id %d", loss);
    int num_sms = 0;
    RBL_TRY(metrics_device(device, &num_sms));
    RBL_CUDA(cudaSetDevice(device));
    return rbl_k_test_metrics(num_sms, X, n, d, ld, w, y, group, loss, threshold, out16, scratch, S(stream));
}

// ---- data ingest on the device (the step before the path) ------------------------------------------------------
int rbl_standardize_scratch_bytes(int device, int64_t ld, int64_t* bytes) {
    RBL_REQUIRE(bytes != nullptr && ld > 0 && ld % 2 == 0, "bad argument (ld must be even)");
    int num_sms = 0;
    RBL_TRY(metrics_device(device, &num_sms));
    int64_t doubles = 0;
    rbl_k_standardize_scratch_doubles(num_sms, ld, &doubles);
    *bytes = doubles * (int64_t)sizeof(double);
    return RBL_OK;
}

int rbl_standardize_columns(int device, double* X, int64_t n, int64_t d, int64_t ld, double* mean_out,
                            double* scale_out, void* scratch, rbl_stream_t stream) {
    RBL_REQUIRE(X && mean_out && scale_out && scratch, "null argument");
    RBL_REQUIRE(n > 0 && d > 0 && ld >= d && ld % 2 == 0, "bad shape: n=%lld d=%lld ld=%lld (ld must be even)",
                (long long)n, (long long)d, (long long)ld);
    RBL_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(mean_out) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(scale_out) % 16 == 0,
                "X, mean_out and scale_out must be 16-byte aligned");
    int num_sms = 0;
    RBL_TRY(metrics_device(device, &num_sms));
    RBL_CUDA(cudaSetDevice(device));
    return rbl_k_standardize(num_sms, X, n, ld, mean_out, scale_out, reinterpret_cast<double*>(scratch), S(stream));
}

int rbl_gather_rows(int device, const double* X, int64_t ld_in, const int64_t* idx, int64_t n_out, int64_t d,
                    double* out, int64_t ld_out, rbl_stream_t stream) {
    RBL_REQUIRE(X && idx && out, "null argument");
    RBL_REQUIRE(n_out > 0 && d > 0 && ld_in >= d && ld_out >= d && ld_in % 2 == 0 && ld_out % 2 == 0,
                "bad shape: n_out=%lld d=%lld ld_in=%lld ld_out=%lld (leading dimensions must be even)",
                (long long)n_out, (long long)d, (long long)ld_in, (long long)ld_out);
    RBL_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "X and out must be 16-byte aligned");
    int num_sms = 0;
    RBL_TRY(metrics_device(device, &num_sms));
    RBL_CUDA(cudaSetDevice(device));
    return rbl_k_gather_rows(num_sms, X, ld_in, idx, n_out, d, out, ld_out, S(stream));
}

}  // extern "C"
