/*
 * rbl_b200.h — C ABI of librbl_b200.so: the B200 (sm_100a) kernels behind the ADMM inner loop of the
 * rank-based-loss framework (SRM / EHRM / AoRR).
 *
 * The reference (RufengXiao/ADMM-for-rank-based-loss) is pure Python and has no FFI; its plug point
 * for this path is the Python class `ADMMmethod` (src/optim/algorithms.py:173-220).  This header is the
 * boundary a native replacement of that class's numerical steps exposes: plain C, opaque handle,
 * `int` status (0 = OK, message via rbl_last_error()), no torch types.  Every entry point cites the
 * reference lines it replaces.  The host layer in admm-for-rank-based-loss_b200/ binds it with ctypes
 * (INTEGRATION.md shows the binding a maintainer of the reference would add).
 *
 * Conventions
 *   - every `const double*` / `double*` / `int32_t*` argument is a DEVICE pointer unless its name starts
 *     with `h_`; the caller (PyTorch) owns all of them; the library owns only handle-internal scratch.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), except rbl_create,
 *     rbl_destroy and rbl_fista_poll (which synchronises the stream to read the device state).
 *   - one handle per (device, problem shape); a handle is not thread-safe, distinct handles are.
 *   - D is row-major n_local x d with leading dimension ld (even, >= d; padding columns must be 0 —
 *     rbl_build_design produces this layout).  Row sharding: this rank holds global rows
 *     [row_lo, row_lo + n_local) of an n_global-row problem; sort and PAV always run on n_global.
 */
#ifndef RBL_B200_H
#define RBL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbl_ctx* rbl_handle_t;
typedef void* rbl_stream_t; /* cudaStream_t */

#define RBL_ABI_VERSION 2

/* loss ids (src/optim/objective.py:26-37 get_loss) */
#define RBL_LOSS_BINARY_CROSS_ENTROPY 0
#define RBL_LOSS_HINGE 1

int rbl_version(void);
const char* rbl_last_error(void);
/* kernels launched through the library by this process (instrumentation for bench.py) */
int64_t rbl_launch_count(void);

/* Per-iteration scalars from DEVICE memory: d_scal = [rho, lam of the FISTA call, thr_f32 (0/1), EHRM clip mode
 * (1 or 2, see rbl_scatter_z)] (4 doubles, caller owned).  While bound (non-NULL), rbl_margins, rbl_pav_prox, rbl_scatter_*, rbl_dual_pass and rbl_gram_fista_run read
 * these instead of their by-value arguments, so the host layer can capture one ADMM iteration as a CUDA graph and
 * replay it after updating d_scal (the rho schedule of algorithms.py:147-157 changes rho every iteration).
 * NULL unbinds. */
int rbl_bind_scalars(rbl_handle_t h, const double* d_scal);

/* handle + scratch.  Replaces the state set up in Optimizer.__init__ (algorithms.py:20-75). */
int rbl_create(rbl_handle_t* out, int device, int64_t n_local, int64_t n_global, int64_t row_lo, int32_t d,
               int64_t ld);
int rbl_destroy(rbl_handle_t h);

/* Persistent CTAs of the D-reading kernels of this handle (default and maximum: one per SM).  Batched mode gives every
 * instance a fraction of the machine so that one instance's HBM-bound gather overlaps the latency-bound sort / PAV /
 * FISTA kernels of the others.  Changes the number of per-CTA partials, i.e. the summation order (rounding only). */
int rbl_set_pass_grid(rbl_handle_t h, int grid);

/* OPTIONAL fp32 storage of the design matrix (north_star: "1e-5 in an optional fp32 mode"; the reference ships a
 * float32 FISTA, algorithms.py:199-201, fast_lasso.py:22-26).  elem_bytes = 4: every `D` / `Dt` / `D_rows` pointer
 * passed to this handle afterwards points to FLOAT rows (leading dimension a multiple of 4 elements = 16 bytes);
 * rbl_build_design* round the fp64 product -y*x once to float.  All dots, column sums, G = D^T D and every vector
 * stay fp64 — only the HBM bytes of the D-reading kernels halve.  elem_bytes = 8 (default) restores fp64.  Call it
 * right after rbl_create, before D is built.  Not available together with rbl_batch_create. */
int rbl_set_storage(rbl_handle_t h, int elem_bytes);
/* h_out[0..8) = num_sms, pass_grid, rows_per_tile, pass_stages, pass_smem_bytes, scratch_bytes, vec_grid,
 * pav_chunk */
int rbl_info(rbl_handle_t h, int64_t* h_out);

/* D = -y (.) X, zero padding up to ld.  algorithms.py:23 */
int rbl_build_design(rbl_handle_t h, const double* X, int64_t ldx, const double* y, double* D, rbl_stream_t stream);

/* the same for `nrows` rows only (X, y, D point at the first of them): lets the host layer build D chunk by chunk
 * while the next chunk of X is still in flight over PCIe */
int rbl_build_design_rows(rbl_handle_t h, const double* X, int64_t ldx, const double* y, double* D, int64_t nrows,
                          rbl_stream_t stream);

/* rank-order spectrum sigma (n_global) used by the PAV: alphas, or betas for EHRM.  algorithms.py:74-75 */
int rbl_set_spectrum(rbl_handle_t h, const double* sigma, rbl_stream_t stream);

/* out = D x  (one pass over D).  algorithms.py:89,132,135 */
int rbl_matvec(rbl_handle_t h, const double* D, const double* x, double* out, rbl_stream_t stream);

/* m = Dw - lambda / rho from a maintained Dw.  algorithms.py:89 */
int rbl_margins(rbl_handle_t h, const double* Dw, const double* lam, double rho, double* m, rbl_stream_t stream);

/* stable ascending key-index sort of n_global margins.  algorithms.py:92-93 */
int rbl_sort_margins(rbl_handle_t h, const double* m, double* m_sorted, int32_t* perm, rbl_stream_t stream);
/* The same sort with a hint: prev_perm = the permutation an EARLIER call returned on similar data (e.g. the previous
 * ADMM iteration; may alias perm, may be NULL; n_global entries).  The rows that sat at B - 1 evenly spaced ranks
 * then, read at their new values and sorted, split the new keys into B ~ n/2048 buckets, each sorted by (key,
 * index) in shared memory by one CTA — two launches, no grid barrier, the same bit-exact stable permutation.  If a
 * bucket would overflow its 4096 slots (stale or useless hint) the LSD sort runs instead, chosen on the device.
 * A hint in a buffer of the caller's own may hold anything (only speed depends on it).  When prev_perm is the very
 * buffer the handle's LAST sort wrote its permutation to (the in-place call of the ADMM loop: prev_perm == perm), it is
 * taken to still hold that permutation untouched, and the keys are partitioned in that rank order (see
 * rbl_sort_config bit 2). */
int rbl_sort_margins_near(rbl_handle_t h, const double* m, const int32_t* prev_perm, double* m_sorted,
                          int32_t* perm, rbl_stream_t stream);
/* synchronises `stream`; h_out[0..4) = buckets B of the splitter sort for this n (0: not used), route the last
 * rbl_sort_margins_near call took (1 buckets, 2 LSD fallback, 0 none yet), its largest bucket, overflow flag now */
int rbl_sort_stats(rbl_handle_t h, rbl_stream_t stream, int32_t* h_out);
/* bit 0: three launches per radix pass instead of the single persistent cooperative kernel; bit 1: never use the
 * splitter sort (testing); bit 2: the splitter sort always partitions in row order with the full splitter search
 * (otherwise, when the hint is the permutation the handle's own last sort wrote, it walks the keys in the previous
 * RANK order: the bucket of a key is then guessed from its previous rank and verified with two loads) */
int rbl_sort_config(rbl_handle_t h, int legacy);
/* dev tool: d_stamps (64 x u64, device) receives %globaltimer stamps of CTA 0 — the persistent sort writes 6 phase
 * boundaries for each of its 8 passes ([pass][6]), the persistent Gram-FISTA kernel a count in [0] followed by its
 * phase boundaries (scripts/sort_phases.py, scripts/fista_phases.py); NULL switches it off */
int rbl_sort_debug(rbl_handle_t h, uint64_t* d_stamps);

/* z_sorted = argmin_{z1<=..<=zn} sum sigma_i loss(z_i) + rho/2 (z_i - m_i)^2.  pav.py:54-178,
 * individual_solver.py:90-130, PAV_cpt.py:169-293 (with sigma = betas; clip in rbl_scatter_z) */
int rbl_pav_prox(rbl_handle_t h, int loss, const double* m_sorted, double rho, double* z_sorted, rbl_stream_t stream);

/* The element prox is already isotonic over every run of ranks where sigma does not increase, so a spectrum with
 * few such runs (ERM 1, superquantile / AoRR <= 3) needs (#runs - 1) merges instead of the merge tree; rbl_set_spectrum
 * finds the runs; the merge searches are warm-started from the pooled blocks the previous rbl_pav_prox call on this
 * handle found (ranks move little between ADMM iterations; the result never depends on the guess).  force_tree
 * bit 0 makes rbl_pav_prox take the general tree anyway, bit 1 switches the warm start off (both for testing);
 * *h_nseg (may be NULL) = number of runs, 0 if there are too many for the few-segment path. */
int rbl_pav_config(rbl_handle_t h, int force_tree, int32_t* h_nseg);

/* out[i] = argmin_z sigma_i loss(z) + rho/2 (z - m_i)^2, no pooling.  individual_solver.py:112-130 */
int rbl_prox_elementwise(rbl_handle_t h, int loss, const double* sigma, const double* m, int64_t n, double rho,
                         double* out, rbl_stream_t stream);

/* EHRM candidate selection, PAV_cpt.py:203-226: out2 (device) = [fval1, fval2] with
 *   fval1 = func_value(sigma_a, min(prox_{sigma_a}(m), B)),  fval2 = func_value(sigma_b, max(prox_{sigma_b}(m), B))
 * at element level on the sorted margins (func_value = sum sigma log(1+e^x) + rho/2 |x - m|^2, :41-43).  The
 * reference takes candidate 1 everywhere when fval1 <= fval2 and candidate 2 otherwise (a scalar comparison); the
 * caller then runs rbl_pav_prox on the chosen spectrum and scatters with use_clip = 2 or 1.  Fixed-order sums. */
int rbl_ehrm_candidate_sums(rbl_handle_t h, const double* m_sorted, const double* sigma_a, const double* sigma_b,
                            double B, double rho, double* out2, rbl_stream_t stream);

/* z[perm] = clipped z_sorted on the rows this rank owns; b = z + lambda/rho (may be NULL).  use_clip: 0 none,
 * 1 max(clip, .) (EHRM candidate 2, PAV_cpt.py:213,271), 2 min(clip, .) (candidate 1, :207,264).  While a scalar
 * block is bound (rbl_bind_scalars) a non-zero use_clip is replaced by (int)scal[3].  algorithms.py:103-104,192 */
int rbl_scatter_z(rbl_handle_t h, const double* z_sorted, const int32_t* perm, int use_clip, double clip,
                  const double* lam, double rho, double* z, double* b, rbl_stream_t stream);

/* rbl_scatter_z that also records the ACTIVE rows: those whose z differs from the margin m.  b - D w = z - m is
 * exactly zero on every other row (sigma_i = 0 outside the pooled blocks; hinge margins below the kink), so the
 * gradient pass of the w-step only has to read the active rows of D.  The list (local row, z - m) is kept in
 * the handle in rank order (deterministic).  algorithms.py:103-104,192 */
int rbl_scatter_active(rbl_handle_t h, const double* z_sorted, const double* m_sorted, const int32_t* perm,
                       int use_clip, double clip, const double* lam, double rho, double* z, double* b,
                       rbl_stream_t stream);
/* red = [D^T (b - D w0) (d), ||b - D w0||^2] at the point w0 the last rbl_scatter_active was taken at
 * (b = z + lambda/rho, m = D w0 - lambda/rho): a gather over the active rows when there are at most `dense_above`
 * of them, else one streaming pass (chosen on the device; r is scratch for the latter).  fast_lasso.py:41-43,
 * w_LBFGS.py:31-45 */
int rbl_grad_pass(rbl_handle_t h, const double* D, const double* w0, const double* b, double* r, int64_t dense_above,
                  double* red, rbl_stream_t stream);
/* the gather kernel of rbl_grad_pass alone, on the current active-row list (per-CTA partials are left in
 * handle scratch): lets bench.py time that kernel by itself */
int rbl_gather_only(rbl_handle_t h, const double* D, rbl_stream_t stream);
/* synchronises `stream`; *h_count = active rows of the last rbl_scatter_active (instrumentation) */
int rbl_active_count(rbl_handle_t h, int32_t* h_count, rbl_stream_t stream);

/* r = b - D x ; red[0..d) = D^T r ; red[d] = ||r||^2  (one pass over D).  w_LBFGS.py:31-45,
 * fast_lasso.py:41-43.  red must hold d + 2 doubles. */
int rbl_fused_pass(rbl_handle_t h, const double* D, const double* x, const double* b, double* r, double* red,
                   rbl_stream_t stream);

/* FISTA for 0.5||b - D beta||^2 + lam ||beta||_1 as a device-resident state machine.
 * fast_lasso.py:22-69 as called from algorithms.py:199-201.
 *   h_pow_tab[i] = float32(eta)**i, i < 128 (host pointer), L0 = float32(17) in the reference. */
int rbl_fista_config(rbl_handle_t h, const float* h_pow_tab);
int rbl_fista_begin(rbl_handle_t h, const double* w0, double lam, int thr_f32, float L0, double tol, int max_iter,
                    rbl_stream_t stream);
/* caller-owned reduce buffer (d + 2 doubles) the passes leave [D^T r, ||r||^2, c0] in, so a row-sharded
 * host layer can all-reduce it in place between rbl_fista_pass and rbl_fista_update (NULL: internal) */
int rbl_fista_bind_red(rbl_handle_t h, double* red);
/* one pass at the current trial point + fixed-order reduction of the per-CTA partials into red */
int rbl_fista_pass(rbl_handle_t h, const double* D, const double* b, rbl_stream_t stream);
int rbl_fista_update(rbl_handle_t h, rbl_stream_t stream);
/* single-GPU convenience: enqueue nsteps x (pass, update) without host round trips */
int rbl_fista_steps(rbl_handle_t h, const double* D, const double* b, int nsteps, rbl_stream_t stream);
/* synchronises `stream`; h_int[0..6) = done, k, passes, trials, i_k, cur; h_dbl[0..4) = crit, L, t, ss */
int rbl_fista_poll(rbl_handle_t h, rbl_stream_t stream, int32_t* h_int, double* h_dbl);
/* w_out (d) = beta, r_out (n_local) = b - D beta of the accepted iterate; either may be NULL */
int rbl_fista_result(rbl_handle_t h, double* w_out, double* r_out, rbl_stream_t stream);

/* ---- Gram mode: the w-step on G = D^T D (d x d) instead of on D.  algorithms.py:24 builds DTD; w_LBFGS.py:39-45
 * uses it for the l2 gradient.  With w0 the warm start and red0 = [g0 = D^T(b - D w0) (d), ss0 = ||b - D w0||^2]
 * from ONE rbl_fused_pass at w0, every FISTA trial / L-BFGS evaluation is a sweep over G (L2-resident) instead
 * of a pass over D, so an ADMM iteration reads D exactly twice (here and in rbl_dual_pass).
 * G is row-major d x ld (same leading dimension as D), exactly symmetric. */
/* G = D^T D over this rank's rows (FP64 tensor cores); a row-sharded host layer all-reduces it once */
int rbl_gram_build(rbl_handle_t h, const double* D, double* G, rbl_stream_t stream);
/* G (+)= D_rows^T D_rows over `nrows` rows starting at D_rows (accumulate = 0 overwrites G): the chunked form of
 * rbl_gram_build, overlapped with the upload of the design matrix */
int rbl_gram_accumulate(rbl_handle_t h, const double* D_rows, int64_t nrows, int accumulate, double* G,
                        rbl_stream_t stream);
/* FISTA (fast_lasso.py:22-69) on G: same trial points, same accept/reject rule (LHS - RHS = D.G D - L||D||^2),
 * same float32 L schedule; state polled with rbl_fista_poll (passes = sweeps over G).  w0 and red0 are
 * caller-owned and must stay valid until the call converges. */
int rbl_gram_fista_begin(rbl_handle_t h, const double* G, const double* w0, const double* red0, double lam,
                         int thr_f32, float L0, double tol, int max_iter, rbl_stream_t stream);
int rbl_gram_fista_steps(rbl_handle_t h, const double* G, int nsteps, rbl_stream_t stream);
/* the whole call as ONE persistent cooperative kernel (G rows and the d-vector state live in shared memory,
 * one grid barrier per trial, no host round trip): runs to convergence, leaves w_out (may be NULL) and the
 * final state for rbl_fista_poll.  rbl_gram_fista_persistent_ok() == 0 (large d): use begin/steps instead. */
int rbl_gram_fista_persistent_ok(rbl_handle_t h);
/* w_out may alias w0 (it is written after the last grid barrier); w_prev_out (may be NULL) receives w0;
 * with_support != 0 also leaves the ascending support of the result in the handle for rbl_dual_pass(support_ready) */
int rbl_gram_fista_run(rbl_handle_t h, const double* G, const double* w0, const double* red0, double lam, int thr_f32,
                       float L0, double tol, int max_iter, double* w_out, double* w_prev_out, int with_support,
                       rbl_stream_t stream);
int rbl_gram_fista_result(rbl_handle_t h, double* w_out, rbl_stream_t stream);
/* red_out = [D^T (b - D w) (d), ||b - D w||^2, 0] at any w, from G, w0, red0 (same layout as rbl_fused_pass):
 * the f/g evaluation of w_LBFGS.py:31-45 */
/* The smooth w-step as ONE call (w_LBFGS.py:48-62: scipy.optimize.minimize(method='L-BFGS-B', maxiter)): the
 * library's own L-BFGS-B (csrc/lbfgs_core.h: scipy's unconstrained control flow — maxcor 10, ftol 2.2e-9, gtol 1e-5,
 * maxls 20, More'-Thuente line search — pinned against scipy in tests/test_host.py) over f/g evaluations on G.
 * reg_kind 0: R(w) = reg/2 ||w||^2 (wl2_fun, :31-45); 1: the Huber-type smoothing of reg/2 ||w||_1 with parameter t
 * (wl1_fun_smooth, :11-28).  h_w (host, d doubles): start on entry, solution on exit, also copied to d_w_out (device,
 * may be NULL) on `stream`; h_info = [iterations, evaluations, status (0 converged, 1 limit, 2 abnormal)].
 * w0 / red0 as in rbl_gram_eval.  Synchronises `stream` once per evaluation. */
int rbl_lbfgs_gram(rbl_handle_t h, const double* G, const double* w0, const double* red0, double rho, double reg,
                   int reg_kind, double t, int maxiter, double* h_w, double* d_w_out, int32_t* h_info,
                   rbl_stream_t stream);

/* rbl_gram_eval for a host-side optimiser (scipy's L-BFGS-B, w_LBFGS.py:48-62): h_w (d doubles) and h_red_out
 * (d + 2 doubles) are HOST arrays; the call stages w through pinned memory, runs the sweep over G, reads the
 * result back and synchronises `stream` — one call per f/g evaluation instead of a copy, a launch, a copy and a
 * synchronisation driven from the interpreter. */
int rbl_gram_eval_host(rbl_handle_t h, const double* G, const double* w0, const double* red0, const double* h_w,
                       double* h_red_out, rbl_stream_t stream);
int rbl_gram_eval(rbl_handle_t h, const double* G, const double* w0, const double* red0, const double* w,
                  double* red_out, rbl_stream_t stream);

/* Dw = D w with the dual update in the epilogue: lambda += rho (z - Dw);
 * out8 = [||z - Dw||^2 (local rows), ||w - w_prev||^2, ||w||^2, ||w||_1, nnz(w), sparse path taken (0/1),
 * FISTA iterations and sweeps of the last w-step, active rows of the last rbl_scatter_active] (9 doubles).
 * algorithms.py:132-136.  The l1 w-step leaves exact zeros in w: when nnz(w) <= sparse_cap the product reads only
 * the nnz(w) touched columns (from the transposed copy Dt when given, else the touched 32-byte sectors of each row of
 * D) — chosen on the device, no host round trip — otherwise one streaming pass over D.  sparse_cap = 0 forces the
 * dense pass. */
/* support_ready != 0: the support of w was left in the handle by rbl_gram_fista_run(with_support).  out8 and
 * w_copy (may be NULL; receives w) may point to pinned host memory (unified addressing): the residuals and the
 * iterate then reach the host without a copy node. */
int rbl_dual_pass(rbl_handle_t h, const double* D, const double* Dt, const double* w, const double* w_prev,
                  const double* z, double* Dw, double* lam, double rho, int sparse_cap, int support_ready,
                  double* out8, double* w_copy, rbl_stream_t stream);
/* Dt = D^T as a dense d x n_local row-major copy (optional, 8 n d bytes).  With it the sparse-w branch of
 * rbl_dual_pass reads the nnz(w) touched columns as contiguous n-vectors (coalesced) instead of one 32-byte sector
 * per row and column (DRAM-activate bound).  Pass Dt = NULL to rbl_dual_pass when no copy is kept. */
int rbl_build_transpose(rbl_handle_t h, const double* D, double* Dt, rbl_stream_t stream);

/* Small-problem l1 w-step (algorithms.py:194-197: n <= 500 and d <= 60 go to sklearn.linear_model.Lasso(alpha,
 * tol=1e-8, fit_intercept=False, max_iter=50000)): scikit-learn's cyclic coordinate descent with its duality-gap
 * stop, from w = 0, for  1/2 ||b - D w||^2 + l1 ||w||_1  (l1 = alpha * n), run on G = D^T D and
 * red0 = [D^T (b - D w_ref) (d), ||b - D w_ref||^2] (the warm-start pass at any point w_ref) — no pass over D.
 * d <= 64.  info3 (device, may be NULL) = [sweeps, duality gap, tol * b.b]. */
int rbl_lasso_cd_gram(rbl_handle_t h, const double* G, const double* w_ref, const double* red0, double l1, double tol,
                      int max_iter, double* w_out, double* info3, rbl_stream_t stream);

/* Upload from PAGEABLE host memory (what a caller's numpy array is): nthreads (<= 16) host threads stage 8 MB chunks
 * through per-thread pinned slots and issue the DMA of each chunk on their own streams; `stream` is made to wait
 * for all of them.  Replaces the single-threaded staging copy the driver performs for a pageable cudaMemcpyAsync.
 * The ingest side of Optimizer.__init__ (algorithms.py:21-23: X.copy(), -y * X). */
int rbl_h2d_pageable(int device, void* d_dst, const void* h_src, int64_t bytes, int nthreads, rbl_stream_t stream);

/* HOST function (no device work): the CPT spectra of EHRM, objective.py:148-164 — which = 0: a_i (gamma 0.69),
 * which = 1: b_i (gamma 0.61), i = 0..n-1 in ascending rank, into the host array h_out.  Scalar libm pow in the
 * reference's order, so the values equal the reference's Python-float loop bit for bit (the differences
 * distort((i+1)/n) - distort(i/n) cancel ~log10(n) digits; a vectorised pow is 3e-9 off at n = 4M). */
int rbl_cpt_weights(int64_t n, int which, double* h_out);

/* ---- native outer loop.  The host layer captures ONE ADMM iteration (z-step, FISTA w-step, dual step, read-back of
 * out8/out9 into pinned memory) as a CUDA graph whose first node copies the pinned scalar block h_scal =
 * [rho, lam, thr_f32] to the block bound with rbl_bind_scalars.  rbl_admm_run then replays it up to max_iters
 * times with the reference's host logic in between (algorithms.py:137-157): stop test ||z - Dw|| < tol and
 * ||w - w_prev|| < tol, rho <- min(rho * (1.02 if primal > 1e-2 else 1.07), 217 d), lam = (reg / (2 rho n)) n.
 * One cudaGraphLaunch + one stream synchronise per iteration, no interpreter in the loop. */
typedef struct rbl_run_stats {
    double rho;             /* rho to use for the NEXT iteration (unchanged by a converged iteration) */
    double primal, dual;    /* residual norms of the last iteration run */
    int32_t iters;          /* iterations run */
    int32_t converged;      /* stop test met at the last iteration */
    int32_t rho_is_pyfloat; /* still the caller's python-float rho (no update happened) */
    int32_t nnz_last;       /* nnz(w) after the last iteration */
    int32_t last_sweeps;    /* sweeps over G of the last FISTA call */
    int32_t pad;
    int64_t fista_iters, fista_sweeps;   /* totals over the run */
    int64_t sparse_dual, dense_dual;     /* which dual-pass branch ran */
    int64_t gathered, rows_read;         /* gradient passes that gathered active rows; rows of D read in total */
} rbl_run_stats;
int rbl_admm_run(rbl_handle_t h, void* graph_exec /* cudaGraphExec_t */, rbl_stream_t stream, double* h_scal,
                 const double* h_out, int32_t max_iters, double tol, double reg, int64_t num_row,
                 int32_t num_feature, int64_t dense_above, double rho, int32_t rho_is_pyfloat, rbl_run_stats* out);

/* The same loop for the l2 problems (w_flag == 2, algorithms.py:109-116 inside :119-157): graph_pre replays the z-step
 * and the warm-start gradient pass (h_w, pinned, receives the warm start; red0 the pass result), the library's
 * L-BFGS-B (rbl_lbfgs_gram, reg_kind 0) solves the w-step, graph_dual replays the dual pass and the read-back into
 * h_out; both graphs start by copying h_scal to the bound scalar block.  In `out`, fista_iters / fista_sweeps count
 * L-BFGS iterations / evaluations.  No interpreter in the loop; not for EHRM (its candidate choice needs the host
 * between the sort and the prox). */
int rbl_admm_run_l2(rbl_handle_t h, void* graph_pre, void* graph_dual, rbl_stream_t stream, double* h_scal,
                    const double* h_out, const double* G, const double* w0, const double* red0, double* h_w,
                    double* d_w, double reg, int32_t lbfgs_maxiter, int32_t max_iters, double tol, int32_t num_feature,
                    int64_t dense_above, double rho, rbl_run_stats* out);

/* ---- batched mode: B independent instances (lambda grid, seeds) sharing one D.  No reference counterpart
 * (the reference runs one ADMMmethod object per instance); SURVEY.md K10.  Buffers are laid out [B][...].
 * One pass over D serves 8 instances at a time (multi-RHS fused pass on the FP64 tensor-core path). */
int rbl_batch_create(rbl_handle_t h, int B);
int rbl_fista_batch_begin(rbl_handle_t h, int B, const double* w0s /* [B][d] */, const double* h_lams /* host [B] */,
                          const int32_t* h_thr_f32 /* host [B] */, float L0, double tol, int max_iter,
                          rbl_stream_t stream);
int rbl_fista_batch_steps(rbl_handle_t h, int B, const double* D, const double* bs /* [B][n_local] */, int nsteps,
                          rbl_stream_t stream);
/* synchronises `stream`; per instance: done flag, outer iterations, passes, final L (h_L may be NULL) */
int rbl_fista_batch_poll(rbl_handle_t h, int B, rbl_stream_t stream, int32_t* h_done, int32_t* h_iters,
                         int32_t* h_passes, double* h_L);
int rbl_fista_batch_result(rbl_handle_t h, int B, double* w_out /* [B][d] */, double* r_out /* [B][n_local] */,
                           rbl_stream_t stream);

/* lambda += rho (z - Dw); out4 = [||z - Dw||^2 (local rows), ||w - w_prev||^2, ||w||^2, ||w||_1].
 * from_residual != 0: Dw := b - r first (r from rbl_fista_result), saving the pass.  algorithms.py:132-136 */
int rbl_dual_update(rbl_handle_t h, const double* z, double* Dw, const double* b, const double* r,
                    int from_residual, double* lam, double rho, const double* w, const double* w_prev, double* out4,
                    rbl_stream_t stream);

/* out4[0] = sum_i sigma_i loss(u_(i)) over the ascending margins u = D w (n_global, unsorted input),
 * out4[2] = ||w||^2, out4[3] = ||w||_1.  objective.py:71-87 */
int rbl_objective(rbl_handle_t h, int loss, const double* margins, const double* sigma, const double* w,
                  double* out4, rbl_stream_t stream);

/* ---- test-set metrics: the step after the path (no handle — a test set has its own row count) ----------------
 * One pass over X (row-major n x d, leading dimension ld) yields every number that
 * src/util/calculate_acc.py:3-19 (calculate_accuracy) and src/util/fair_metric.py:3-40 (calculate_statistics)
 * derive their results from:
 *   out16[0]  rows whose prediction equals the label (BCE: +1 iff sigmoid(x.w) >= threshold, else -1;
 *             hinge: always +1, as the reference ships it, calculate_acc.py:14-15)
 *   out16[1]  n
 *   out16[2 + 6 g + k], group g in {0, 1}:  k = 0 rows, 1 predicted positive, 2 TP, 3 FN, 4 TN, 5 FP
 *   out16[14] sum b,  out16[15] sum b log b,  b = sigmoid(x.w) - y01 + 1     (Theil index, fair_metric.py:35-38)
 * y holds the labels as doubles (-1 / +1); group (int32, may be NULL: all rows in group 0).  Counts are exact
 * integers in doubles; the two sums are reduced in a fixed order.  scratch: rbl_metrics_scratch_bytes() bytes of
 * device memory, zeroed once before the first call (the call leaves it ready for the next). */
int rbl_metrics_scratch_bytes(int device, int64_t* bytes);
int rbl_test_metrics(int device, const double* X, int64_t n, int64_t d, int64_t ld, const double* w,
                     const double* y, const int32_t* group, int loss, double threshold, double* out16,
                     void* scratch, rbl_stream_t stream);

/* ---- data ingest on the device: the step before the path (no handle) ------------------------------------------
 * rbl_standardize_columns: X <- (X - mean) / std per column, in place — `preprocessing.scale(X)` of
 * src/util/load_data.py:115 (scikit-learn semantics: population std, columns with std < 10 eps keep scale 1).
 * X is row-major n x d with an EVEN leading dimension ld (padding columns must be zero and stay zero);
 * mean_out / scale_out receive ld doubles each.  scratch: rbl_standardize_scratch_bytes(device, ld) bytes.
 * rbl_gather_rows: out[i, :] = X[idx[i], :] for i < n_out (int64 row indices on the device) — the row selection of
 * the drivers' train_test_split (run_SRM.py:26). */
int rbl_standardize_scratch_bytes(int device, int64_t ld, int64_t* bytes);
int rbl_standardize_columns(int device, double* X, int64_t n, int64_t d, int64_t ld, double* mean_out,
                            double* scale_out, void* scratch, rbl_stream_t stream);
int rbl_gather_rows(int device, const double* X, int64_t ld_in, const int64_t* idx, int64_t n_out, int64_t d,
                    double* out, int64_t ld_out, rbl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RBL_B200_H */
