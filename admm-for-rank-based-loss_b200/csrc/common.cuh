// common.cuh — internal declarations shared by the translation units of librbl_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "prox_core.h"

#define RBL_OK 0
#define RBL_ERR_CUDA 1
#define RBL_ERR_ARG 2
#define RBL_ERR_UNSUPPORTED 3

void rbl_set_error(const char* fmt, ...);

#define RBL_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            rbl_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
            return RBL_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

extern long long g_rbl_launches;  // kernels launched by this process through the library
extern int g_rbl_pdl;             // programmatic dependent launch for the kernels of the iteration (RBL_PDL=1: on)

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still
// draining, and waits (rbl_pdl_wait, first statement of the kernel) until that predecessor has completed and
// flushed before it touches memory.  Meant to hide the launch latency between the ~15 short kernels of one ADMM
// iteration in eager (non-graph) runs; opt-in, see api.cu.
#if defined(__CUDACC__)
__device__ __forceinline__ void rbl_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t rbl_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                  Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_rbl_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
// cudaFuncSetAttribute is per device: "already raised" flags are kept per device ordinal, so a process that drives
// several devices (handles on cuda:0 and cuda:1) opts every one of them in
#define RBL_MAX_DEVICES 64
#define RBL_PER_DEVICE(type, name, c) \
    static type name##_dev[RBL_MAX_DEVICES] = {}; \
    type& name = name##_dev[(c)->device & (RBL_MAX_DEVICES - 1)]

#define RBL_LAUNCH_CHECK()             \
    do {                               \
        ++g_rbl_launches;              \
        RBL_CUDA(cudaGetLastError());  \
    } while (0)

#define RBL_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            rbl_set_error(__VA_ARGS__);                                                             \
            return RBL_ERR_ARG;                                                                     \
        }                                                                                           \
    } while (0)

// ---- FISTA device-resident state (fast_lasso.py:22-69 control flow, kept on the device so the
// host never has to synchronise inside the inner loop) ---------------------------------------------
struct FistaState {
    double t;        // momentum scalar (fast_lasso.py:39,59)
    double c0;       // ||b - D beta_p||^2  (drbp, :42)
    double rhs;      // L*||beta - beta_p||^2 - 2 (beta - beta_p).g_p  (:53)
    double tol;
    double lam;
    double t1;       // (t-1)/tnext of the last accepted step (:61)
    double crit;     // ||beta_k - beta_{k-1}||  (:63)
    double ss_last;  // ||b - D beta||^2 of the last pass
    float L_prev;    // np.float32 scalars, as the reference passes them (algorithms.py:199-200)
    float L_cur;
    int i_k;         // line-search index (:44-47)
    int k;           // -1 before the initial pass, else outer iteration count
    int max_iter;
    int done;
    int passes;      // D passes executed
    int trials;      // line-search trials executed
    int thr_f32;     // lam/L_cur evaluated in float32 (python-float lam under NEP 50) or float64
    int accepted;    // set by the update kernel for the combine kernel: 0 none, 1 accepted, 2 init
    int cur;         // which residual buffer the next pass writes (the other holds r at beta_prev)
    int c0_from_red; // c0 must be taken from red[d+1] (written by the combine kernel)
    int pad0, pad1;
};

// internal context behind rbl_handle_t
struct rbl_ctx {
    int device;
    int num_sms;
    int64_t n_local, n_global, row_lo;
    int d;
    int64_t ld;
    int esz;            // bytes per stored element of D / D^T: 8 (fp64, default) or 4 (optional fp32 storage)
    // ---- pass kernels
    int pass_grid;      // persistent CTAs
    int pass_rows;      // rows per tile
    int pass_stages;
    size_t pass_smem;
    double* gpart;      // [pass_grid][ld] per-CTA column partials
    double* sspart;     // [pass_grid]
    // ---- reductions of n-vectors
    int vec_grid;
    double* vpart;      // [4][vec_grid]
    // ---- FISTA
    FistaState* fista;      // device
    FistaState* fista_host; // pinned mirror
    float* pow_tab;         // device, 128 entries eta**i as float32
    double *beta, *beta_p, *beta_prev, *g_p, *g_prev;  // d-vectors (ld padded)
    double* rbuf[2];        // n_local residual buffers
    double* red;            // [d + 2]: g (d), ss, c0 — internal (red_own) or bound by the host layer
    double* red_own;
    double* c0part;         // [vec_grid]
    // ---- Gram-mode w-step (gram_kernels.cu): q(beta_prev), right-hand sides, products, last-CTA ticket
    double *gq_prev, *gxs, *gvu;
    unsigned int* gticket;
    double* gvu2;                       // [2][8][ld] parity-double-buffered products of the persistent kernel
    int gp_checked, gp_grid, gp_rpc, gp_g_in_smem, gp_kc;  // persistent FISTA kernel shape (gp_grid = 0: unavailable)
    size_t gp_smem;
    const double *gram_w0, *gram_red0;  // caller-owned warm start and [g0, ss0] of the running FISTA call
    // ---- support of w (sparse D w in the dual pass): ascending column indices, values, count
    int32_t* sup_idx;
    double* sup_val;
    int* sup_nnz;
    // ---- active rows of the last rbl_scatter_active: rows whose z differs from the margin (rank order)
    int* act_cta_count;   // [vec_grid]
    int32_t* act_row;     // [n_local] local row indices
    double* act_delta;    // [n_local] z - m
    int* act_total;
    // ---- sort
    uint64_t *keysA, *keysB;
    uint32_t *valsA, *valsB;
    uint32_t* tile_hist;    // [256 * ntiles] per-digit tile counts, then [256] digit totals
    int sort_tiles;
    uint32_t* sort_counts;  // [num_sms][256] per-CTA digit counts of the persistent sort
    int psort_checked, psort_ok, sort_legacy;
    // splitter sort (sort_kernels.cu): bucket slots, counts, overflow flag; ss_nb = 0: not used for this n
    int ss_nb, ss_off, ss_row_order;
    const int32_t* last_perm;  // the permutation buffer the last sort on this handle wrote (and its length):
    int64_t last_perm_n;       // a hint equal to it is known to be a permutation (rank-order partition)
    uint64_t *ss_bkey, *ss_spl;
    uint32_t *ss_bval, *ss_count;
    int* ss_flag;
    unsigned long long* sort_dbg;  // dev tool: phase timestamps of the persistent sort (null: off)
    // ---- PAV
    int chunk_log2;
    int64_t nchunks;
    double *ps_loc_hi, *ps_loc_lo, *ps_off_hi, *ps_off_lo;  // sigma prefixes (set once)
    double *ps_tot_hi, *ps_tot_lo;                          // per-chunk totals of sigma
    double *pm_loc_hi, *pm_loc_lo, *pm_off_hi, *pm_off_lo;  // margin prefixes (every z-step)
    double *ch_tot_hi, *ch_tot_lo;                          // per-chunk totals scratch
    unsigned int* node_cnt;  // merge-tree arrival counters (self-cleaning)
    double* sigma;          // n_global, rank order (the sigma the PAV uses: alphas, or betas for EHRM)
    double* val;            // n_global block values
    const double* scal;     // caller-owned device scalars [rho, lam_fista, thr_f32] (rbl_bind_scalars) or null
    int has_sigma;
    int nseg;               // runs of non-increasing sigma (0: too many, use the merge tree)
    int force_tree;         // testing: always take the merge tree
    int pav_no_hints;       // testing: few-segment merge without the warm start from the previous call's blocks
    int* seg_count;
    int64_t* seg_bounds;    // [nseg + 1] device
    void* seg_blocks;       // SegBlocks (pooled blocks of the last few-segment call)
    // ---- batched mode (K10): B instances sharing D
    int batch_cap;          // instances the batched buffers were sized for (0: not created)
    int batch_stages;
    FistaState* bfista;     // [B] device
    FistaState* bfista_host;  // [2B] pinned: poll mirror, begin staging
    double *bbeta, *bbeta_p, *bbeta_prev, *bg_p, *bg_prev;  // [B][ld+8]
    double* brbuf[2];       // [B][n_local]
    double* bred;           // [B][ld+8]
    double *bgpart, *bsspart, *bc0part;  // [grid][8][ld], [grid][8], [B][vec_grid]
    // ---- host-driven f/g evaluations (rbl_gram_eval_host): pinned staging and device copies, created on first use
    double *eval_stage, *eval_w, *eval_red;
    // ---- objective
    double* obj_tmp;        // n_global
    size_t bytes;           // total scratch allocated
    char* slab;             // all handle scratch lives in one allocation (api.cu: ctx_alloc_slab)
    size_t slab_bytes, slab_off;
    int alloc_mode;         // 0 direct cudaMalloc, 1 measuring, 2 carving
    void* extra[32];        // direct allocations made after creation (batched mode)
    int n_extra;
};

// ---- launchers implemented in the kernel translation units --------------------------------------
int rbl_launch_pass(rbl_ctx* c, int mode, const double* D, const double* x, const double* b, double* out,
                    const FistaState* st, double* const* rbuf, cudaStream_t s, double* lam = nullptr,
                    double rho = 0.0, const int* gate_nnz = nullptr, int gate_cap = 0);
int rbl_pass_configure(rbl_ctx* c);
// column partials of sum_k delta_k D[row_k] over the active-row list (runs iff *count <= cap)
int rbl_launch_gather(rbl_ctx* c, const double* D, const int32_t* rows, const double* delta, const int* count,
                      int cap, cudaStream_t s);

#define RBL_PASS_MATVEC 0  // out = D x
#define RBL_PASS_FUSED 1   // out = r = b - D x ; column partials of D^T r ; partial ||r||^2
#define RBL_PASS_FISTA 2   // as FUSED with x = ctx->beta, out = rbuf[st->cur], skipped when st->done
#define RBL_PASS_DUAL 3    // out = D x ; lam += rho (b - D x) with b = z ; partial ||z - D x||^2
