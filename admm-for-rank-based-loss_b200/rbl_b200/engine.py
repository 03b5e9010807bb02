"""Host-side engine: device buffers (torch), streams, row sharding (torch.distributed) and the calls
into librbl_b200.so that make up one ADMM iteration.

PyTorch is plumbing here (device memory, streams, NCCL); every numerical step is a hand-written
sm_100a kernel reached through the C ABI (include/rbl_b200.h).  There is no CPU fallback: without a
B200 and the built library, construction raises.

Reference call stack this replaces: Optimizer.__init__ / z_subproblem / main_loop
(src/optim/algorithms.py:20-164) and ADMMmethod._w_subproblem (:190-207).
"""
import ctypes
import os

import numpy as np
import torch

from . import _cabi

LOSS_IDS = {"binary_cross_entropy": 0, "hinge": 1}

_POW_CACHE = {}


def _pow_table(eta):
    """float32(eta)**i for i < 128, computed by numpy exactly as fast_lasso.py:46 does (eta**i_k)."""
    key = float(eta)
    if key not in _POW_CACHE:
        with np.errstate(over="ignore"):
            _POW_CACHE[key] = np.array([np.float32(eta) ** i for i in range(128)], dtype=np.float32)
    return _POW_CACHE[key]


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise _cabi.RblError("no CUDA device: rbl_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def shard_bounds(n, world, rank):
    """Contiguous row shards, sizes differing by at most one (first n % world ranks get the extra row)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class DeviceProblem:
    """D = -y (.) X resident in HBM (row-major, padded to an even leading dimension) plus the library
    handle that owns the scratch for passes, sort and PAV over it."""

    def __init__(self, X, y, device=None, group=None, row_lo=None, n_global=None, _share=None, _want_gram=False,
                 storage=None):
        """storage: "fp64" (default) or "fp32" — the OPTIONAL mode that keeps D (and D^T) in float32 in HBM: every
        D-reading kernel moves half the bytes; all arithmetic (dots, column sums, G = D^T D, every vector) stays
        fp64.  Results then differ from the fp64 run by the rounding of D's entries (~1e-7 relative)."""
        self.lib = _cabi.load()
        self._parent = _share
        storage = (storage or os.environ.get("RBL_STORAGE", "fp64")).lower()
        if storage in ("float32", "f32"):
            storage = "fp32"
        if storage in ("float64", "f64"):
            storage = "fp64"
        if storage not in ("fp64", "fp32"):
            raise ValueError(f"storage must be 'fp64' or 'fp32' (got {storage!r})")
        self.storage = storage if _share is None else _share.storage
        self.esz = 4 if self.storage == "fp32" else 8
        self.store_dtype = torch.float32 if self.esz == 4 else torch.float64
        if _share is not None:
            # another solver instance over the SAME design matrix (batched mode): own handle, scratch and state,
            # D / G / D^T borrowed from the parent
            p = _share
            self.device, self.group, self.world, self.rank = p.device, p.group, p.world, p.rank
            self.n_local, self.d, self.n_global, self.row_lo, self.ld = p.n_local, p.d, p.n_global, p.row_lo, p.ld
            with torch.cuda.device(self.device):
                h = ctypes.c_void_p()
                _cabi.check(self.lib.rbl_create(ctypes.byref(h), self.device.index or 0, self.n_local, self.n_global,
                                                self.row_lo, self.d, self.ld))
                self.h = h
                if self.esz == 4:
                    _cabi.check(self.lib.rbl_set_storage(self.h, 4))
                self.D = p.D
                self._out4 = torch.zeros(16, dtype=torch.float64, device=self.device)
                self._out4_host = torch.zeros(16, dtype=torch.float64).pin_memory()
                self._out4_np = self._out4_host.numpy()
                self._u_local = torch.empty(self.n_local, dtype=torch.float64, device=self.device)
                self._u_glob = (torch.empty(self.n_global, dtype=torch.float64, device=self.device)
                                if self.world > 1 else self._u_local)
            self.info = dict(p.info)
            return
        self.device = _require_cuda(device)
        self.group = group
        if row_lo is None:  # unsharded: this process holds every row
            row_lo, n_global = 0, X.shape[0]
            self.world, self.rank = 1, 0
        else:               # X holds global rows [row_lo, row_lo + n_local) of an n_global-row problem
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
            # the all-gather of the margins (gather_rows) assumes THIS partition on every rank
            want = shard_bounds(int(n_global), self.world, self.rank)
            if (int(row_lo), int(row_lo) + int(X.shape[0])) != want:
                raise ValueError(f"rank {self.rank} of {self.world} must hold rows [{want[0]}, {want[1]}) of the "
                                 f"{int(n_global)}-row problem (rbl_b200.engine.shard_bounds), got "
                                 f"[{int(row_lo)}, {int(row_lo) + int(X.shape[0])})")
        self.n_local, self.d = int(X.shape[0]), int(X.shape[1])
        self.n_global, self.row_lo = int(n_global), int(row_lo)
        self.ld = self.d + (self.d & 1) if self.esz == 8 else (self.d + 3) // 4 * 4   # 16-byte rows
        import time as _time
        self.build_times = {}
        with torch.cuda.device(self.device):
            t0 = _time.perf_counter()
            h = ctypes.c_void_p()
            _cabi.check(self.lib.rbl_create(ctypes.byref(h), self.device.index or 0, self.n_local, self.n_global,
                                            self.row_lo, self.d, self.ld))
            self.h = h
            if self.esz == 4:
                _cabi.check(self.lib.rbl_set_storage(self.h, 4))
            t1 = _time.perf_counter()
            self.G = None
            yd = self._to_device(np.asarray(y, dtype=np.float64).reshape(-1) if not torch.is_tensor(y)
                                 else y.reshape(-1).to(torch.float64))
            self.D = torch.empty((self.n_local, self.ld), dtype=self.store_dtype, device=self.device)
            chunk = int(os.environ.get("RBL_PIPELINE_ROWS", "0")) or max(4096, (256 << 20) // (8 * self.d))  # ~256 MB
            if (not torch.is_tensor(X)) and self.n_local >= 4 * chunk and os.environ.get("RBL_PIPELINE", "1") != "0":
                # host array: upload X chunk by chunk on a copy stream while the previous chunk is turned into rows
                # of D = -y*X and (Gram mode) accumulated into G = D^T D — the 2 n d^2 flop of G hide behind PCIe
                self._upload_pipelined(X, yd, chunk, _want_gram)
                t2 = t3 = _time.perf_counter()
            else:
                Xd = self._to_device(X).reshape(self.n_local, self.d)
                torch.cuda.current_stream().synchronize()
                t2 = _time.perf_counter()
                _cabi.check(self.lib.rbl_build_design(self.h, Xd.data_ptr(), self.d, yd.data_ptr(),
                                                      self.D.data_ptr(), self._stream()))
                torch.cuda.current_stream().synchronize()
                t3 = _time.perf_counter()
                del Xd
            del yd
            self.build_times.update(handle_scratch_s=t1 - t0, h2d_s=t2 - t1, design_s=t3 - t2)
            self._out4 = torch.zeros(16, dtype=torch.float64, device=self.device)
            self._out4_host = torch.zeros(16, dtype=torch.float64).pin_memory()
            self._out4_np = self._out4_host.numpy()
            self._u_local = torch.empty(self.n_local, dtype=torch.float64, device=self.device)
            self._u_glob = (torch.empty(self.n_global, dtype=torch.float64, device=self.device)
                            if self.world > 1 else self._u_local)
        info = (ctypes.c_int64 * 8)()
        _cabi.check(self.lib.rbl_info(self.h, info))
        self.info = dict(zip(["num_sms", "pass_grid", "rows_per_tile", "pass_stages", "pass_smem", "scratch_bytes",
                              "vec_grid", "pav_chunk"], list(info)))

    def _upload_pipelined(self, X, yd, chunk, want_gram):
        n, d, dev = self.n_local, self.d, self.device
        Xt = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).reshape(n, d)  # no copy for fp64 C-order
        main = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty((chunk, d), dtype=torch.float64, device=dev) for _ in range(2)]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        if want_gram:
            self.G = torch.empty((d, self.ld), dtype=torch.float64, device=dev)
        copy_stream.wait_stream(main)
        ms = ctypes.c_void_p(main.cuda_stream)
        # a caller's numpy array is PAGEABLE: the driver would stage it through one small pinned buffer on this
        # thread (~10 GB/s); librbl_b200 stages it with several host threads through pinned slots instead
        pageable = (not Xt.is_pinned()) and os.environ.get("RBL_PAGEABLE_STAGING", "1") != "0"
        threads = int(os.environ.get("RBL_UPLOAD_THREADS", "0")) or min(8, max(1, (os.cpu_count() or 2) // 2))
        self.upload_path = f"pageable: {threads} staging threads (rbl_h2d_pageable)" if pageable else "pinned: direct DMA"
        cs = ctypes.c_void_p(copy_stream.cuda_stream)
        for k, r0 in enumerate(range(0, n, chunk)):
            r1, b = min(n, r0 + chunk), k % 2
            with torch.cuda.stream(copy_stream):
                if k >= 2:
                    copy_stream.wait_event(consumed[b])
                if pageable:
                    _cabi.check(self.lib.rbl_h2d_pageable(dev.index or 0, stage[b].data_ptr(),
                                                          Xt[r0:r1].data_ptr(), (r1 - r0) * d * 8, threads, cs))
                else:
                    stage[b][: r1 - r0].copy_(Xt[r0:r1], non_blocking=True)
                copied[b].record(copy_stream)
            main.wait_event(copied[b])
            Dk = self.D[r0:r1]
            _cabi.check(self.lib.rbl_build_design_rows(self.h, stage[b].data_ptr(), d, yd[r0:].data_ptr(),
                                                       Dk.data_ptr(), r1 - r0, ms))
            if want_gram:
                _cabi.check(self.lib.rbl_gram_accumulate(self.h, Dk.data_ptr(), r1 - r0, 1 if k else 0,
                                                         self.G.data_ptr(), ms))
            consumed[b].record(main)
        main.synchronize()
        self.gram_during_upload = bool(want_gram)

    # ---- plumbing ----------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _to_device(self, a):
        if torch.is_tensor(a):
            return a.to(device=self.device, dtype=torch.float64).contiguous()
        a = np.ascontiguousarray(a, dtype=np.float64)
        return torch.from_numpy(a).to(self.device)

    def vec(self, a):
        """d- or n-vector to a contiguous fp64 device tensor"""
        if torch.is_tensor(a):
            return a.to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
        return self._to_device(np.asarray(a, dtype=np.float64).reshape(-1))

    @property
    def launches(self):
        """kernels launched through librbl_b200 by this process (bench.py's gpu_launches)"""
        return (int(self.lib.rbl_launch_count())
                + getattr(self, "_graph_replays", 0) * getattr(self, "_graph_launches", 0)
                + getattr(self, "_graph_l2_replays", 0) * getattr(self, "_graph_l2_launches", 0)
                + getattr(self, "_graph_dual_replays", 0) * getattr(self, "_graph_dual_launches", 0))

    def close(self):
        # a captured iteration graph holds NCCL work in a row-sharded job: drop it before the handle (and before
        # the caller destroys the process group)
        if (getattr(self, "_graph", None) is not None or getattr(self, "_graph_l2", None) is not None
                or getattr(self, "_graph_dual", None) is not None):
            torch.cuda.synchronize(self.device)
            self._graph = self._graph_l2 = self._graph_dual = None
            self._graphs_l2_ehrm = {}
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.rbl_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def gather_rows(self, local, out):
        """all-gather of an n_local vector into the n_global vector `out` (NCCL over NVLink)."""
        if self.world == 1:
            if out.data_ptr() != local.data_ptr():
                out.copy_(local)
            return out
        dist = torch.distributed
        base, extra = divmod(self.n_global, self.world)
        if extra == 0:
            dist.all_gather_into_tensor(out, local, group=self.group)
        else:
            pad = base + 1
            buf = torch.zeros(self.world * pad, dtype=local.dtype, device=self.device)
            mine = torch.zeros(pad, dtype=local.dtype, device=self.device)
            mine[: self.n_local] = local
            dist.all_gather_into_tensor(buf, mine, group=self.group)
            for r in range(self.world):
                lo, hi = shard_bounds(self.n_global, self.world, r)
                out[lo:hi] = buf[r * pad: r * pad + (hi - lo)]
        return out

    def all_reduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t, group=self.group)
        return t

    # ---- kernels -----------------------------------------------------------------------------
    def matvec(self, x, out=None):
        """out = D x  (algorithms.py:89,132,135)"""
        if out is None:
            out = torch.empty(self.n_local, dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.rbl_matvec(self.h, self.D.data_ptr(), x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def objective_terms(self, w, sigma, loss, u_local=None):
        """(sum_i sigma_i loss(u_(i)), ||w||^2, ||w||_1) for u = D w  (objective.py:71-87).
        u_local: the margins D w of this rank's rows when the caller already holds them (no pass over D)"""
        if u_local is None:
            self.matvec(w, self._u_local)
            u_local = self._u_local
        self.gather_rows(u_local, self._u_glob)
        _cabi.check(self.lib.rbl_objective(self.h, LOSS_IDS[loss], self._u_glob.data_ptr(), sigma.data_ptr(),
                                           w.data_ptr(), self._out4.data_ptr(), self._stream()))
        self._out4_host.copy_(self._out4, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        o = self._out4_host
        return float(o[0]), float(o[2]), float(o[3])


class AdmmEngine(DeviceProblem):
    """ADMM state and the three sub-steps on the device."""

    @classmethod
    def child(cls, parent):
        """a further solver instance over the parent's design matrix (shares D, G = D^T D and D^T)"""
        return cls(None, None, parent.loss, parent.sigma, clip=parent.clip, ehrm=parent.ehrm, _share=parent)

    def __init__(self, X, y, loss, sigma, clip=None, ehrm=None, **kw):
        """sigma: the spectrum of the isotonic prox; clip: lower clip B applied after it (EHRM candidate 2 only —
        kept for callers that fix the candidate); ehrm=(sigma_a, sigma_b, B): the reference's EHRM z-step with its
        per-call choice between min(B, prox_{sigma_a}) and max(B, prox_{sigma_b}) (PAV_cpt.py:203-226); `sigma` and
        `clip` are then ignored."""
        import time as _time
        _t0 = _time.perf_counter()
        if kw.get("_share") is None and X is not None:
            # the w-step formulation is known from the shape alone: let the upload accumulate G on the way
            n_g, d_ = int(kw.get("n_global") or np.shape(X)[0]), int(np.shape(X)[1])
            m_ = os.environ.get("RBL_W_MODE", "auto").lower()
            kw["_want_gram"] = (m_ == "gram") or (m_ == "auto" and d_ <= 4096 and n_g >= 2 * d_)
        super().__init__(X, y, **kw)
        _t1 = _time.perf_counter()
        if loss not in LOSS_IDS:
            raise ValueError(f"Unrecognized loss '{loss}'! Options: ['binary_cross_entropy', 'hinge']")
        self.loss, self.loss_id = loss, LOSS_IDS[loss]
        self.ehrm = ehrm
        if ehrm is not None:
            if loss != "binary_cross_entropy":
                raise ValueError("ehrm only can be with binary_cross_entropy.")  # pav.py:60
            sigma, clip = ehrm[1], float(ehrm[2])  # candidate 2 until the first z-step has compared the two sums
        self.clip = clip
        self.clip_mode = 0 if clip is None else 1   # rbl_scatter_z use_clip: 1 = max(B, .), 2 = min(B, .)
        self.ehrm_stats = {"cand1": 0, "cand2": 0, "switches": 0, "last": None}
        dev, f64 = self.device, torch.float64
        nl, ng, d = self.n_local, self.n_global, self.d
        with torch.cuda.device(dev):
            self.sigma = self.vec(sigma)
            assert self.sigma.numel() == ng
            if ehrm is not None:
                self.sig_a, self.sig_b = self.vec(ehrm[0]), self.sigma
                assert self.sig_a.numel() == ng
                self._ehrm_dev = torch.zeros(2, dtype=f64, device=dev)
                self._ehrm_host = torch.zeros(2, dtype=f64).pin_memory()
            _cabi.check(self.lib.rbl_set_spectrum(self.h, self.sigma.data_ptr(), self._stream()))
            self.w = torch.zeros(d, dtype=f64, device=dev)
            self.w_prev = torch.zeros(d, dtype=f64, device=dev)
            self.z = torch.zeros(nl, dtype=f64, device=dev)
            self.lam = torch.zeros(nl, dtype=f64, device=dev)
            self.Dw = torch.zeros(nl, dtype=f64, device=dev)
            self.m = torch.zeros(nl, dtype=f64, device=dev)
            self.b = torch.zeros(nl, dtype=f64, device=dev)
            self.r = torch.zeros(nl, dtype=f64, device=dev)
            self.m_glob = torch.zeros(ng, dtype=f64, device=dev) if self.world > 1 else self.m
            self.m_sorted = torch.zeros(ng, dtype=f64, device=dev)
            self.z_sorted = torch.zeros(ng, dtype=f64, device=dev)
            self.perm = torch.zeros(ng, dtype=torch.int32, device=dev)
            self.red = torch.zeros(d + 2, dtype=f64, device=dev)
            self.red_host = torch.zeros(d + 2, dtype=f64).pin_memory()
            self.w_host = torch.zeros(d, dtype=f64).pin_memory()
            _cabi.check(self.lib.rbl_fista_bind_red(self.h, self.red.data_ptr()))
        self.Dw_valid = False
        self.fista_stats = {"calls": 0, "passes": 0, "iters": 0, "polls": 0, "last_passes": 0, "d_passes": 0}
        # w-step formulation: "gram" runs FISTA / L-BFGS f-g on G = D^T D (two passes over D per ADMM
        # iteration), "stream" runs every trial as a pass over D.  auto: gram whenever G is the smaller object.
        mode = os.environ.get("RBL_W_MODE", "auto").lower()
        if mode not in ("auto", "gram", "stream"):
            raise ValueError(f"RBL_W_MODE must be auto, gram or stream (got {mode!r})")
        if mode == "auto":
            mode = "gram" if (d <= 4096 and ng >= 2 * d) else "stream"
        self.w_mode = mode
        if not hasattr(self, "G"):
            self.G = None
        self.gram_build_s = 0.0
        self._persistent = None
        self._fista_info_pending = False
        self._sorted_valid = False
        self.splitter_sort = os.environ.get("RBL_SPLITTER_SORT", "1") != "0"
        # gradient pass: gather over the active rows (z != m) unless more than this fraction of rows is active
        # (the gather moves its bytes at 1.01 of the copy peak, the streaming pass at 1.05: break-even ~0.96)
        self.active_dense_frac = float(os.environ.get("RBL_ACTIVE_FRAC", "0.9"))
        self._delta_valid = False
        self._active_pending = False
        self.active_stats = {"calls": 0, "rows": 0, "gathered": 0}
        # CUDA-graph replay of the whole iteration (iteration_fista)
        self.graph_ok = os.environ.get("RBL_GRAPH", "1") != "0"
        self.graph_mgpu = os.environ.get("RBL_GRAPH_MGPU", "1") != "0"  # capture the NCCL collectives as well
        self._graph, self._graph_key, self._iters_eager = None, None, 0
        self._graph_l2, self._graph_l2_active, self._iters_eager_l2, self._pre_done = None, False, 0, False
        self._graphs_l2_ehrm = {}  # EHRM: one graph of (PAV, scatter, gradient pass) per candidate
        # dual pass: D w reads only the touched sectors of D when nnz(w) <= sparse_cap (0 disables)
        self.sparse_cap = int(os.environ.get("RBL_SPARSE_CAP", str(max(1, d // 16))))
        self.dual_stats = {"sparse": 0, "dense": 0, "nnz_last": d}
        # transposed copy of D for the sparse-w dual pass: built once w has come out sparse, if it fits
        self.Dt = None
        self.transpose_ok = os.environ.get("RBL_TRANSPOSE", "1") != "0"
        torch.cuda.current_stream(dev).synchronize()
        _t2 = _time.perf_counter()
        if mode == "gram":
            self.gram()  # like the reference, which forms DTD in Optimizer.__init__ (algorithms.py:24)
        if hasattr(self, "build_times"):
            self.build_times.update(problem_total_s=_t1 - _t0, state_and_spectrum_s=_t2 - _t1,
                                    gram_total_s=_time.perf_counter() - _t2)
        self._fista_eta = None
        self.fista_batch_min = int(os.environ.get("RBL_FISTA_BATCH", "4"))

    # ---- state -------------------------------------------------------------------------------
    def set_state(self, w=None, z=None, lam=None):
        self._delta_valid = False
        self._support_ready = False
        self._iters_eager = 0  # a few eager iterations again before the graph path resumes
        self._iters_eager_l2, self._pre_done = 0, False
        if w is not None:
            self.w.copy_(self.vec(w))
            self.Dw_valid = False
        if z is not None:
            self.z.copy_(self.vec(z))
        if lam is not None:
            self.lam.copy_(self.vec(lam))

    def refresh_Dw(self):
        self.matvec(self.w, self.Dw)
        self.Dw_valid = True

    # ---- Gram matrix (algorithms.py:24), built on first use ------------------------------------------
    def gram(self):
        if self.G is None and self._parent is not None:
            with torch.cuda.device(self.device):
                self.G = self._parent.gram()
                self.red0 = torch.zeros(self.d + 2, dtype=torch.float64, device=self.device)
                self.red1 = torch.zeros(self.d + 2, dtype=torch.float64, device=self.device)
                _cabi.check(self.lib.rbl_fista_bind_red(self.h, self.red0.data_ptr()))
            self._gram_reduced = True
        if self.G is None:
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.device(self.device):
                self.G = torch.empty((self.d, self.ld), dtype=torch.float64, device=self.device)
                t0.record()
                _cabi.check(self.lib.rbl_gram_build(self.h, self.D.data_ptr(), self.G.data_ptr(), self._stream()))
                t1.record()
                t1.synchronize()
            self.gram_build_s = t0.elapsed_time(t1) * 1e-3
            self._gram_reduced = False
        if not getattr(self, "_gram_reduced", False):
            self.all_reduce(self.G)  # row shards: G is the sum of the per-rank Grams
            self._gram_reduced = True
        if not hasattr(self, "red0"):
            with torch.cuda.device(self.device):
                self.red0 = torch.zeros(self.d + 2, dtype=torch.float64, device=self.device)
                self.red1 = torch.zeros(self.d + 2, dtype=torch.float64, device=self.device)
                # reductions of the gradient pass land in red0 directly (no copy node)
                _cabi.check(self.lib.rbl_fista_bind_red(self.h, self.red0.data_ptr()))
        return self.G

    def _build_transpose(self):
        if self._parent is not None:
            if self._parent.Dt is None and self._parent.transpose_ok:
                self._parent._build_transpose()
            self.Dt, self.transpose_ok = self._parent.Dt, self._parent.transpose_ok
            return
        free, _ = torch.cuda.mem_get_info(self.device)
        need = self.n_local * self.d * self.esz
        if need + (2 << 30) > free:
            self.transpose_ok = False  # not enough HBM left: keep the sector-gather kernel
            return
        with torch.cuda.device(self.device):
            Dt = torch.empty((self.d, self.n_local), dtype=self.store_dtype, device=self.device)
            _cabi.check(self.lib.rbl_build_transpose(self.h, self.D.data_ptr(), Dt.data_ptr(), self._stream()))
            # one-off: finish before publishing it — child engines (batched mode) read it from other streams
            torch.cuda.current_stream(self.device).synchronize()
            self.Dt = Dt

    def _pass_at(self, w0, b, use_active=False):
        """red0 = [D^T (b - D w0), ||b - D w0||^2] — the one pass over D a Gram-mode w-step makes.  Right after
        a z-step, b - D w0 = z - m is zero on every row the prox left alone: only the active rows are read."""
        if use_active:
            _cabi.check(self.lib.rbl_grad_pass(self.h, self.D.data_ptr(), w0.data_ptr(), b.data_ptr(),
                                               self.r.data_ptr(), int(self.active_dense_frac * self.n_local),
                                               self.red0.data_ptr(), self._stream()))
            self._active_pending = True
        else:
            _cabi.check(self.lib.rbl_fused_pass(self.h, self.D.data_ptr(), w0.data_ptr(), b.data_ptr(),
                                                self.r.data_ptr(), self.red0.data_ptr(), self._stream()))
            self.fista_stats["d_passes"] += 1
        self.all_reduce(self.red0)

    # ---- z-step: margins -> sort -> PAV prox -> scatter (algorithms.py:88-106) -----------------
    def z_step(self, rho):
        self._z_pre(rho)
        return self._z_post(rho)

    def _z_pre(self, rho):
        """margins, all-gather, sort; for EHRM also the reference's candidate choice (PAV_cpt.py:203-226): the two
        scalar sums are formed on the device, read back (one 16-byte copy + stream synchronisation) and the
        spectrum / clip direction of the isotonic prox that follows is switched when the winner changes"""
        lib, s = self.lib, self._stream()
        if not self.Dw_valid:
            self.refresh_Dw()
        _cabi.check(lib.rbl_margins(self.h, self.Dw.data_ptr(), self.lam.data_ptr(), float(rho), self.m.data_ptr(), s))
        self.gather_rows(self.m, self.m_glob)
        if self._sorted_valid and self.splitter_sort:
            # the rank order of the previous z-step supplies the splitters of this one
            _cabi.check(lib.rbl_sort_margins_near(self.h, self.m_glob.data_ptr(), self.perm.data_ptr(),
                                                  self.m_sorted.data_ptr(), self.perm.data_ptr(), s))
        else:
            _cabi.check(lib.rbl_sort_margins(self.h, self.m_glob.data_ptr(), self.m_sorted.data_ptr(),
                                             self.perm.data_ptr(), s))
            self._sorted_valid = True
        if self.ehrm is not None:
            self._ehrm_select(rho)

    def _ehrm_select(self, rho):
        lib, s = self.lib, self._stream()
        _cabi.check(lib.rbl_ehrm_candidate_sums(self.h, self.m_sorted.data_ptr(), self.sig_a.data_ptr(),
                                                self.sig_b.data_ptr(), float(self.ehrm[2]), float(rho),
                                                self._ehrm_dev.data_ptr(), s))
        self._ehrm_host.copy_(self._ehrm_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        f1, f2 = float(self._ehrm_host[0]), float(self._ehrm_host[1])
        mode = 2 if f1 <= f2 else 1   # :222-226: opt_array1 everywhere iff fval1 <= fval2, else opt_array2
        st = self.ehrm_stats
        st["cand1" if mode == 2 else "cand2"] += 1
        st["last"] = (f1, f2)
        if mode != self.clip_mode:
            st["switches"] += 1
            self.clip_mode = mode
            self.sigma = self.sig_a if mode == 2 else self.sig_b
            _cabi.check(lib.rbl_set_spectrum(self.h, self.sigma.data_ptr(), s))
        if hasattr(self, "_scal_np"):
            self._scal_np[3] = float(self.clip_mode)

    def _z_post(self, rho):
        lib, s = self.lib, self._stream()
        clip = 0.0 if self.clip is None else float(self.clip)
        _cabi.check(lib.rbl_pav_prox(self.h, self.loss_id, self.m_sorted.data_ptr(), float(rho),
                                     self.z_sorted.data_ptr(), s))
        if self.w_mode == "gram" and self.active_dense_frac > 0:
            _cabi.check(lib.rbl_scatter_active(self.h, self.z_sorted.data_ptr(), self.m_sorted.data_ptr(),
                                               self.perm.data_ptr(), self.clip_mode, clip, self.lam.data_ptr(),
                                               float(rho), self.z.data_ptr(), self.b.data_ptr(), s))
            self._delta_valid = True  # until w, z or lambda change
        else:
            _cabi.check(lib.rbl_scatter_z(self.h, self.z_sorted.data_ptr(), self.perm.data_ptr(), self.clip_mode,
                                          clip, self.lam.data_ptr(), float(rho), self.z.data_ptr(),
                                          self.b.data_ptr(), s))
        return self.z

    # ---- w-step, l1: FISTA (fast_lasso.py:22-69 via algorithms.py:190-202) ---------------------
    def fista(self, w0, b, lam, L=np.float32(17), eta=np.float32(2.5), tol=7e-5, max_iter=5000, w_out=None,
              r_out=None, want_info=True, use_active=False, w_prev_out=None):
        """Runs the device-resident FISTA state machine to completion; returns (w_out, info).

        `lam` keeps its Python type on purpose: a python float makes `lam/L_cur` a float32 quotient
        under numpy >= 2 (as in iteration 0 of the reference loop), np.float64 a float64 one.
        """
        lib, s = self.lib, self._stream()
        if self._fista_eta != float(eta):
            tab = _pow_table(eta)
            _cabi.check(lib.rbl_fista_config(self.h, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
            self._fista_eta = float(eta)
        thr_f32 = 1 if type(lam) is float or isinstance(lam, (int, np.float32)) else 0
        if self.w_mode != "gram":
            _cabi.check(lib.rbl_fista_begin(self.h, w0.data_ptr(), float(lam), thr_f32, float(np.float32(L)),
                                            float(tol), int(max_iter), s))
        hi = (ctypes.c_int32 * 8)()
        hd = (ctypes.c_double * 4)()
        if self.w_mode == "gram":
            return self._fista_gram(w0, b, lam, thr_f32, L, tol, max_iter, w_out, r_out, hi, hd, want_info,
                                    use_active, w_prev_out)
        # first batch: what the previous call needed (iteration counts drift slowly between ADMM
        # iterations), then small batches; steps enqueued after convergence exit immediately
        batch = max(self.fista_batch_min, self.fista_stats["last_passes"] - 1)
        Dp, bp = self.D.data_ptr(), b.data_ptr()
        while True:
            if self.world == 1:
                _cabi.check(lib.rbl_fista_steps(self.h, Dp, bp, batch, s))
            else:
                for _ in range(batch):
                    _cabi.check(lib.rbl_fista_pass(self.h, Dp, bp, s))
                    self.all_reduce(self.red)
                    _cabi.check(lib.rbl_fista_update(self.h, s))
            _cabi.check(lib.rbl_fista_poll(self.h, s, hi, hd))
            self.fista_stats["polls"] += 1
            if hi[0]:
                break
            batch = self.fista_batch_min
        if w_out is None:
            w_out = torch.empty(self.d, dtype=torch.float64, device=self.device)
        _cabi.check(lib.rbl_fista_result(self.h, w_out.data_ptr(), 0 if r_out is None else r_out.data_ptr(), s))
        info = {"iters": int(hi[1]), "passes": int(hi[2]), "trials": int(hi[3]), "L": float(hd[1]),
                "crit": float(hd[0])}
        st = self.fista_stats
        st["calls"] += 1
        st["passes"] += info["passes"]
        st["iters"] += info["iters"]
        st["last_passes"] = info["passes"]
        st["d_passes"] += info["passes"]
        return w_out, info

    def _fista_gram(self, w0, b, lam, thr_f32, L, tol, max_iter, w_out, r_out, hi, hd, want_info=True,
                    use_active=False, w_prev_out=None):
        """FISTA on G = D^T D: one fused pass over D at the warm start, then one sweep over G per trial —
        the whole call is one persistent cooperative kernel when its state fits in shared memory."""
        lib, s = self.lib, self._stream()
        G = self.gram()
        if w_out is None:
            w_out = torch.empty(self.d, dtype=torch.float64, device=self.device)
        if self._persistent is None:
            self._persistent = bool(lib.rbl_gram_fista_persistent_ok(self.h)) and \
                os.environ.get("RBL_GRAM_PERSISTENT", "1") != "0"
        if w0.data_ptr() == w_out.data_ptr() and not self._persistent:
            raise ValueError("w0 and w_out must not alias in Gram mode without the persistent kernel")
        self._pass_at(w0, b, use_active)
        if self._persistent:
            # the kernel also emits w_prev (= w0) and the support of the result for the sparse dual pass
            _cabi.check(lib.rbl_gram_fista_run(self.h, G.data_ptr(), w0.data_ptr(), self.red0.data_ptr(), float(lam),
                                               thr_f32, float(np.float32(L)), float(tol), int(max_iter),
                                               w_out.data_ptr(), 0 if w_prev_out is None else w_prev_out.data_ptr(),
                                               1, s))
            self._support_ready = w_out.data_ptr() == self.w.data_ptr()
            if want_info:
                _cabi.check(lib.rbl_fista_poll(self.h, s, hi, hd))
        else:
            _cabi.check(lib.rbl_gram_fista_begin(self.h, G.data_ptr(), w0.data_ptr(), self.red0.data_ptr(),
                                                 float(lam), thr_f32, float(np.float32(L)), float(tol),
                                                 int(max_iter), s))
            batch = max(self.fista_batch_min, self.fista_stats["last_passes"] + 2)
            while True:
                _cabi.check(lib.rbl_gram_fista_steps(self.h, G.data_ptr(), batch, s))
                _cabi.check(lib.rbl_fista_poll(self.h, s, hi, hd))
                self.fista_stats["polls"] += 1
                if hi[0]:
                    break
                batch = 2 * self.fista_batch_min
            _cabi.check(lib.rbl_gram_fista_result(self.h, w_out.data_ptr(), s))
            want_info = True
        if r_out is not None:  # seam users that want r = b - D beta: one more pass
            self.matvec(w_out, r_out)
            torch.sub(b, r_out, out=r_out)
        st = self.fista_stats
        st["calls"] += 1
        if not want_info:
            self._fista_info_pending = True  # k and sweeps arrive with the dual step's read-back
            return w_out, {"mode": "gram"}
        info = {"iters": int(hi[1]), "passes": int(hi[2]), "trials": int(hi[3]), "L": float(hd[1]),
                "crit": float(hd[0]), "mode": "gram"}
        st["passes"] += info["passes"]
        st["iters"] += info["iters"]
        st["last_passes"] = info["passes"]
        return w_out, info

    def w_step_fista(self, lam, tol=7e-5, max_iter=5000):
        if self.w_mode == "gram":
            use_active, self._delta_valid = self._delta_valid, False
            if self._persistent:  # in place: the kernel reads w, writes the new w and the old one to w_prev
                _, info = self.fista(self.w, self.b, lam, tol=tol, max_iter=max_iter, w_out=self.w,
                                     want_info=False, use_active=use_active, w_prev_out=self.w_prev)
            else:
                self.w_prev.copy_(self.w)
                _, info = self.fista(self.w_prev, self.b, lam, tol=tol, max_iter=max_iter, w_out=self.w,
                                     want_info=False, use_active=use_active)
            self._r_matches_w = False
            return info
        self.w_prev.copy_(self.w)
        _, info = self.fista(self.w_prev, self.b, lam, tol=tol, max_iter=max_iter, w_out=self.w, r_out=self.r)
        self._r_matches_w = True
        return info

    # ---- w-step, l1, small problems: scikit-learn's Lasso coordinate descent (algorithms.py:194-197) --------
    def w_step_lasso_cd(self, l1, tol=1e-8, max_iter=50000):
        """w = argmin 1/2 ||b - D w||^2 + l1 ||w||_1 by sklearn's cyclic coordinate descent (from w = 0, duality-gap
        stop) on G = D^T D: one warm-start pass over D (active rows only right after a z-step) for D^T b and b.b,
        then the whole solve in one single-warp kernel.  l1 = alpha * n with the reference's alpha = reg/(2 rho n)."""
        self.gram()
        if not hasattr(self, "_cd_info"):
            self._cd_info = torch.zeros(4, dtype=torch.float64, device=self.device)
        use_active, self._delta_valid = self._delta_valid, False
        self.w_prev.copy_(self.w)
        self._pass_at(self.w_prev, self.b, use_active)
        _cabi.check(self.lib.rbl_lasso_cd_gram(self.h, self.G.data_ptr(), self.w_prev.data_ptr(),
                                               self.red0.data_ptr(), float(l1), float(tol), int(max_iter),
                                               self.w.data_ptr(), self._cd_info.data_ptr(), self._stream()))
        self._r_matches_w = False
        self._support_ready = False
        return {"mode": "lasso_cd"}

    def lasso_cd_info(self):
        """(sweeps, duality gap, gap tolerance) of the last w_step_lasso_cd (synchronises)"""
        v = self._cd_info.cpu().numpy()
        return int(v[0]), float(v[1]), float(v[2])

    # ---- w-step, l2: host L-BFGS-B over fused device f/g (w_LBFGS.py:31-53) ---------------------
    def fg_smooth(self, w_np, rho, reg_fg):
        """f = rho/2 ||D w - b||^2 + R(w), g = rho D^T(D w - b) + R'(w) — the n x d part is ONE fused pass
        over D on the device; `reg_fg(w) -> (R, R')` is a d-vector formula evaluated on the host."""
        if self.w_mode == "gram":
            # one sweep over G: [D^T (b - D w), ||b - D w||^2] from the pass made at the warm start; host array in,
            # host array out, one library call (staging, launch, read-back, synchronisation)
            w_np = np.ascontiguousarray(w_np, dtype=np.float64)
            red = self._fg_red
            _cabi.check(self.lib.rbl_gram_eval_host(self.h, self.G.data_ptr(), self.w_prev.data_ptr(),
                                                    self.red0.data_ptr(), w_np.ctypes.data, red.ctypes.data,
                                                    self._stream()))
            R, dR = reg_fg(w_np)
            f = 0.5 * rho * float(red[self.d]) + R
            g = -rho * red[: self.d] + dR
            self._last_eval = w_np.copy()
            self.lbfgs_evals += 1
            return f, g
        self._wtmp.copy_(torch.from_numpy(w_np))
        _cabi.check(self.lib.rbl_fused_pass(self.h, self.D.data_ptr(), self._wtmp.data_ptr(), self.b.data_ptr(),
                                            self.r.data_ptr(), self.red.data_ptr(), self._stream()))
        self.all_reduce(self.red)
        self.fista_stats["d_passes"] += 1
        self.red_host.copy_(self.red, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        red = self.red_host.numpy()
        R, dR = reg_fg(w_np)
        f = 0.5 * rho * float(red[self.d]) + R
        g = -rho * red[: self.d] + dR
        self._last_eval = w_np.copy()
        self.lbfgs_evals += 1
        return f, g

    def _warm_start_pass(self):
        """what every L-BFGS w-step starts from: w_prev = w (device and host) and, in Gram mode,
        red0 = [D^T (b - D w), ||b - D w||^2] from the one pass over D (active rows only right after a z-step)"""
        self.w_prev.copy_(self.w)
        self.w_host.copy_(self.w, non_blocking=True)
        if self.w_mode == "gram":
            self.gram()
            use_active, self._delta_valid = self._delta_valid, False
            self._pass_at(self.w_prev, self.b, use_active)

    def z_and_grad(self, rho):
        """z-step + the warm-start pass of an L-BFGS w-step (l2 and smoothed-l1 problems).  In Gram mode the launch
        sequence is static (every data-dependent branch is taken on the device), so after two eager iterations
        it is captured as a CUDA graph and replayed with rho refreshed in the device scalar block: ~20 launches
        become one; the host then only drives scipy's L-BFGS-B over rbl_gram_eval."""
        can_graph = (self.graph_ok and self.w_mode == "gram" and (self.world == 1 or self.graph_mgpu)
                     and self._iters_eager_l2 >= 2 and self.Dw_valid and self.G is not None
                     and not getattr(self, "_r_matches_w", False))
        if can_graph and self.ehrm is not None:
            # the candidate choice needs the host (one 16-byte read-back): margins + sort + the two sums run eagerly,
            # the rest (PAV, scatter, gradient pass) is the replayed graph of the chosen candidate
            self._ensure_scalars()
            self._z_pre(rho)
            self._graph_l2 = self._graphs_l2_ehrm.get(self.clip_mode)
        if can_graph and self._graph_l2 is None:
            self._capture_l2()
            if self.ehrm is not None and self._graph_l2 is not None:
                self._graphs_l2_ehrm[self.clip_mode] = self._graph_l2
        if can_graph and self._graph_l2 is not None:
            self._scal_np[0] = float(rho)
            self._scal_np[3] = float(self.clip_mode)
            self._graph_l2.replay()
            self._graph_l2_replays += 1
            self._delta_valid, self._pre_done = False, True
            self._active_pending = self._graph_l2_active
            return
        self._iters_eager_l2 += 1
        self.z_step(rho)
        self._warm_start_pass()
        self._pre_done = True

    def _ensure_scalars(self):
        if not hasattr(self, "scal"):
            self.scal = torch.zeros(4, dtype=torch.float64, device=self.device)
            self.scal_host = torch.zeros(4, dtype=torch.float64).pin_memory()
            self._scal_np = self.scal_host.numpy()
            self._scal_np[3] = float(self.clip_mode)  # [rho, lam, thr_f32, clip mode]

    def _capture_l2(self):
        dev = self.device
        self._ensure_scalars()
        torch.cuda.current_stream(dev).synchronize()
        stats = (dict(self.fista_stats), dict(self.dual_stats), dict(self.active_stats))
        g = torch.cuda.CUDAGraph()
        n0 = int(self.lib.rbl_launch_count())
        _cabi.check(self.lib.rbl_bind_scalars(self.h, self.scal.data_ptr()))
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_cap_stream"):
            self._cap_stream = torch.cuda.Stream(device=dev)
        try:
            self._cap_stream.wait_stream(cur)
            with torch.cuda.stream(self._cap_stream):
                g.capture_begin()
                try:
                    self.scal.copy_(self.scal_host, non_blocking=True)
                    if self.ehrm is not None:             # (margins, sort and the candidate choice ran eagerly)
                        self._z_post(1.0)
                    else:
                        self.z_step(1.0)                  # by-value scalars are ignored while bound
                    self._graph_l2_active = bool(self._delta_valid)
                    self._warm_start_pass()
                finally:
                    g.capture_end()
            cur.wait_stream(self._cap_stream)
            self._graph_l2 = g
            self._graph_l2_launches = int(self.lib.rbl_launch_count()) - n0
            self._graph_l2_replays = getattr(self, "_graph_l2_replays", 0) if self.ehrm is not None else 0
        except Exception as exc:  # noqa: BLE001 — capture unsupported here: stay on eager launches
            import warnings
            warnings.warn(f"rbl_b200: CUDA graph capture of the z-step + gradient pass failed ({exc!r}); "
                          "continuing with eager launches", RuntimeWarning, stacklevel=2)
            self._graph_l2, self.graph_ok = None, False
        finally:
            _cabi.check(self.lib.rbl_bind_scalars(self.h, 0))
            self.fista_stats, self.dual_stats, self.active_stats = stats
            self._active_pending = False

    # ---- native loop of l2 iterations (rbl_admm_run_l2) ------------------------------------------------------------
    def _capture_dual(self):
        """the dual step (dual pass, [all-reduce of the primal partial], read-back of the residuals and of w into
        pinned memory) as a CUDA graph whose first node refreshes the bound scalar block"""
        dev = self.device
        self._ensure_scalars()
        torch.cuda.current_stream(dev).synchronize()
        stats = (dict(self.fista_stats), dict(self.dual_stats), dict(self.active_stats))
        g = torch.cuda.CUDAGraph()
        n0 = int(self.lib.rbl_launch_count())
        _cabi.check(self.lib.rbl_bind_scalars(self.h, self.scal.data_ptr()))
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_cap_stream"):
            self._cap_stream = torch.cuda.Stream(device=dev)
        try:
            self._cap_stream.wait_stream(cur)
            with torch.cuda.stream(self._cap_stream):
                g.capture_begin()
                try:
                    self.scal.copy_(self.scal_host, non_blocking=True)
                    self._r_matches_w = False
                    self._dual_launch(1.0)                # by-value rho is ignored while the block is bound
                finally:
                    g.capture_end()
            cur.wait_stream(self._cap_stream)
            self._graph_dual = g
            self._graph_dual_launches = int(self.lib.rbl_launch_count()) - n0
        except Exception as exc:  # noqa: BLE001
            import warnings
            warnings.warn(f"rbl_b200: CUDA graph capture of the dual step failed ({exc!r}); continuing with the "
                          "per-iteration loop", RuntimeWarning, stacklevel=2)
            self._graph_dual, self.graph_ok = None, False
        finally:
            _cabi.check(self.lib.rbl_bind_scalars(self.h, 0))
            self.fista_stats, self.dual_stats, self.active_stats = stats
            self._active_pending = False

    def run_lbfgs_iterations(self, n_iters, rho, reg, num_feature, tol, lbfgs_maxiter=1000):
        """up to n_iters ADMM iterations of an l2 problem (stopping at the reference's stop test) in the library's
        native loop: graph (z-step + gradient pass) -> L-BFGS-B in the library -> graph (dual step); returns the
        rbl_run_stats, or None while the graphs are not available (first iterations, EHRM, stream mode)"""
        ok = (n_iters > 0 and self.graph_ok and self.w_mode == "gram" and self.ehrm is None
              and (self.world == 1 or self.graph_mgpu) and self._iters_eager_l2 >= 2 and self.Dw_valid
              and self.G is not None and not getattr(self, "_r_matches_w", False)
              and os.environ.get("RBL_LBFGS", "native") != "scipy")
        if not ok:
            return None
        if self._graph_l2 is None:
            self._capture_l2()
        if self._graph_l2 is not None and getattr(self, "_graph_dual", None) is None:
            self._capture_dual()
        if self._graph_l2 is None or getattr(self, "_graph_dual", None) is None:
            return None
        if not hasattr(self, "_lb_info"):
            self._lb_info = (ctypes.c_int32 * 4)()
            self.lbfgs_evals = 0
        st = _cabi.RunStats()
        self._pre_done = False
        _cabi.check(self.lib.rbl_admm_run_l2(
            self.h, ctypes.c_void_p(self._graph_l2.raw_cuda_graph_exec()),
            ctypes.c_void_p(self._graph_dual.raw_cuda_graph_exec()), self._stream(),
            ctypes.c_void_p(self.scal_host.data_ptr()), ctypes.c_void_p(self._out4_host.data_ptr()),
            self.G.data_ptr(), self.w_prev.data_ptr(), self.red0.data_ptr(), ctypes.c_void_p(self.w_host.data_ptr()),
            self.w.data_ptr(), float(reg), int(lbfgs_maxiter), int(n_iters), float(tol), int(num_feature),
            int(self.active_dense_frac * self.n_local), float(rho), ctypes.byref(st)))
        self._graph_l2_replays += st.iters
        self._graph_dual_replays = getattr(self, "_graph_dual_replays", 0) + st.iters
        self.Dw_valid, self._delta_valid, self._r_matches_w = True, False, False
        self.lbfgs_evals += st.fista_sweeps
        ds, a = self.dual_stats, self.active_stats
        ds["sparse"] += st.sparse_dual
        ds["dense"] += st.dense_dual
        ds["nnz_last"] = st.nnz_last
        a["calls"] += st.iters
        a["rows"] += st.rows_read
        a["gathered"] += st.gathered
        self.fista_stats["d_passes"] += (st.iters - st.gathered) + st.dense_dual
        return st

    def w_step_lbfgs(self, rho, reg, maxiter=1000, reg_fg=None, huber_t=None):
        """The smooth w-step (w_LBFGS.py:48-62).  huber_t=None: the l2 problem (:31-53); huber_t=t: the smoothed-l1
        problem of smoothADMMmethod (:11-28, 54-62).
        Gram mode: ONE library call — librbl_b200's own L-BFGS-B (scipy's unconstrained control flow, pinned against
        scipy in tests/test_host.py) over device f/g evaluations on G; no interpreter inside the w-step
        (RBL_LBFGS=scipy keeps scipy's minimize over rbl_gram_eval_host for cross-checks).
        Stream mode: scipy's L-BFGS-B over the fused pass; `reg_fg(w) -> (R, R')` may replace the regulariser."""
        if not hasattr(self, "_wtmp"):
            self._wtmp = torch.zeros(self.d, dtype=torch.float64, device=self.device)
            self._fg_red = np.zeros(self.d + 2, dtype=np.float64)
            self._lb_info = (ctypes.c_int32 * 4)()
            self.lbfgs_evals = 0
        pre_done, self._pre_done = getattr(self, "_pre_done", False), False
        if not pre_done:  # (z_and_grad has already made the warm-start pass, inside its graph)
            self._warm_start_pass()
        rho, reg = float(rho), float(reg)
        native = (self.w_mode == "gram" and reg_fg is None and os.environ.get("RBL_LBFGS", "native") != "scipy")
        if native:
            torch.cuda.current_stream(self.device).synchronize()
            wh = self.w_host.numpy()          # pinned copy of the warm start, overwritten with the solution
            _cabi.check(self.lib.rbl_lbfgs_gram(self.h, self.G.data_ptr(), self.w_prev.data_ptr(),
                                                self.red0.data_ptr(), rho, reg, 0 if huber_t is None else 1,
                                                0.0 if huber_t is None else float(huber_t), int(maxiter),
                                                wh.ctypes.data, self.w.data_ptr(), self._lb_info, self._stream()))
            self.lbfgs_evals += int(self._lb_info[1])
            self._r_matches_w = False
            return {"nit": int(self._lb_info[0]), "nfev": int(self._lb_info[1]), "solver": "librbl_b200"}
        from scipy.optimize import minimize

        torch.cuda.current_stream(self.device).synchronize()
        w0 = self.w_host.numpy().copy()
        if reg_fg is None and huber_t is not None:
            t = float(huber_t)

            def reg_fg(w):  # wl1_fun_smooth / wl1_fun_smooth_gradient, w_LBFGS.py:11-28
                small = np.abs(w) <= t
                R = 0.5 * 0.5 * reg * float(np.sum(np.square(w[small]))) / t
                R += 0.5 * reg * float(np.sum(np.abs(w[~small]) - 0.5 * t))
                return R, np.where(small, 0.5 * reg * w / t, 0.5 * reg * np.sign(w))
        if reg_fg is None:
            def reg_fg(w):  # wl2_fun / wl2_fun_gradient regulariser, w_LBFGS.py:36,44
                return 0.5 * reg * float(w @ w), reg * w
        res = minimize(lambda w: self.fg_smooth(w, rho, reg_fg), w0, jac=True, method="L-BFGS-B",
                       options={"maxiter": maxiter})
        self.w.copy_(torch.from_numpy(res.x))
        # L-BFGS-B normally returns the last point it evaluated; then r = b - D w is already there
        self._r_matches_w = self.w_mode != "gram" and bool(np.array_equal(res.x, self._last_eval))
        return {"nit": int(res.nit), "nfev": int(res.nfev), "solver": "scipy"}

    # ---- dual update + residuals (algorithms.py:132-136) -----------------------------------------
    def _dual_launch(self, rho):
        """enqueue the dual update and the read-back of its 128-byte result; no host synchronisation"""
        from_res = 1 if getattr(self, "_r_matches_w", False) else 0
        direct = not from_res and self.world == 1  # residuals and w written straight into pinned host memory
        if not from_res:
            # Dw = D w with the multiplier update and ||z - Dw||^2 in the pass epilogue
            sup_ready, self._support_ready = (1 if getattr(self, "_support_ready", False) else 0), False
            _cabi.check(self.lib.rbl_dual_pass(self.h, self.D.data_ptr(),
                                               0 if self.Dt is None else self.Dt.data_ptr(), self.w.data_ptr(),
                                               self.w_prev.data_ptr(),
                                               self.z.data_ptr(), self.Dw.data_ptr(), self.lam.data_ptr(),
                                               float(rho), self.sparse_cap, sup_ready,
                                               self._out4_host.data_ptr() if direct else self._out4.data_ptr(),
                                               self.w_host.data_ptr() if direct else 0, self._stream()))
        else:
            _cabi.check(self.lib.rbl_dual_update(self.h, self.z.data_ptr(), self.Dw.data_ptr(), self.b.data_ptr(),
                                                 self.r.data_ptr(), from_res, self.lam.data_ptr(), float(rho),
                                                 self.w.data_ptr(), self.w_prev.data_ptr(), self._out4.data_ptr(),
                                                 self._stream()))
        self.Dw_valid = True
        self._r_matches_w = False
        if direct:
            return from_res
        if self.world > 1:
            # only the primal term is a partial sum over row shards
            self.all_reduce(self._out4[:1])
        self._out4_host.copy_(self._out4, non_blocking=True)
        self.w_host.copy_(self.w, non_blocking=True)
        return from_res

    def _dual_finish(self, from_res):
        """after the stream has been synchronised: bookkeeping from the read-back, residual norms"""
        o = self._out4_np  # numpy view of the pinned read-back buffer
        if not from_res:
            took_sparse = o[5] != 0.0
            ds = self.dual_stats
            ds["sparse" if took_sparse else "dense"] += 1
            ds["nnz_last"] = int(o[4])
            st = self.fista_stats
            if self._active_pending:
                self._active_pending = False
                rows = int(o[8])
                gathered = rows <= int(self.active_dense_frac * self.n_local)
                a = self.active_stats
                a["calls"] += 1
                a["rows"] += rows if gathered else self.n_local
                a["gathered"] += 1 if gathered else 0
                st["d_passes"] += 0 if gathered else 1
            if self._fista_info_pending:
                self._fista_info_pending = False
                st["iters"] += int(o[6])
                st["passes"] += int(o[7])
                st["last_passes"] = int(o[7])
            st["d_passes"] += 0 if took_sparse else 1
        return float(np.sqrt(o[0])), float(np.sqrt(o[1]))

    def dual_step(self, rho):
        """lambda += rho (z - D w); returns (||z - D w||_2, ||w - w_prev||_2)."""
        if (self.Dt is None and self.dual_stats["sparse"] >= 1 and self.transpose_ok
                and not getattr(self, "_r_matches_w", False)):
            self._build_transpose()
        from_res = self._dual_launch(rho)
        torch.cuda.current_stream(self.device).synchronize()
        return self._dual_finish(from_res)

    # ---- one whole ADMM iteration with the FISTA w-step, replayed as a CUDA graph ------------------------
    def iteration_fista(self, rho, lam, tol=7e-5, max_iter=5000, sync=True):
        """z-step, l1 w-step, dual step (algorithms.py:119-136).  Every data-dependent branch of the Gram-mode
        iteration is taken on the device, so the launch sequence is static: after a few eager iterations it is
        captured once as a CUDA graph and replayed with the per-iteration scalars (rho, lam) refreshed in a
        device block (rbl_bind_scalars) — ~20 launches become one.  Returns (primal, dual); with sync=False the
        work is only enqueued on the current stream (batched mode: many instances in flight) and the caller
        synchronises that stream before finish_iteration()."""
        thr_f32 = 1.0 if type(lam) is float or isinstance(lam, (int, np.float32)) else 0.0
        key = (float(tol), int(max_iter))
        can_graph = (self.graph_ok and self.w_mode == "gram" and (self.world == 1 or self.graph_mgpu)
                     and self._persistent is True and self.ehrm is None
                     and self._iters_eager >= 2 and self.Dw_valid and (self.Dt is not None or not self.transpose_ok
                                                     or self.dual_stats["sparse"] < 1)
                     and not getattr(self, "_r_matches_w", False))
        if can_graph and (self._graph is None or self._graph_key != key):
            self._capture_iteration(key)
        if can_graph and self._graph is not None:
            sh = self._scal_np
            sh[0], sh[1], sh[2] = float(rho), float(lam), thr_f32
            self._graph.replay()
            self._graph_replays += 1
            self.Dw_valid, self._delta_valid, self._r_matches_w = True, False, False
            self._active_pending = self._fista_info_pending = True
            self.fista_stats["calls"] += 1
            self._inflight = 0
            if not sync:
                return None
            torch.cuda.current_stream(self.device).synchronize()
            return self._dual_finish(0)
        self._iters_eager += 1
        self.z_step(rho)
        self.w_step_fista(lam, tol=tol, max_iter=max_iter)
        if (self.Dt is None and self.dual_stats["sparse"] >= 1 and self.transpose_ok
                and not getattr(self, "_r_matches_w", False)):
            self._build_transpose()
        self._inflight = self._dual_launch(rho)
        if not sync:
            return None
        torch.cuda.current_stream(self.device).synchronize()
        return self._dual_finish(self._inflight)

    def finish_iteration(self):
        """after the stream of an iteration_fista(sync=False) call has been synchronised: (primal, dual)"""
        return self._dual_finish(self._inflight)

    def graph_ready(self, tol=7e-5, max_iter=5000):
        """True once the iteration graph exists (capturing it now if the engine state allows it)"""
        key = (float(tol), int(max_iter))
        ok = (self.graph_ok and self.w_mode == "gram" and (self.world == 1 or self.graph_mgpu)
              and self._persistent is True and self.ehrm is None and self._iters_eager >= 2 and self.Dw_valid
              and (self.Dt is not None or not self.transpose_ok or self.dual_stats["sparse"] < 1)
              and not getattr(self, "_r_matches_w", False))
        if ok and (self._graph is None or self._graph_key != key):
            self._capture_iteration(key)
        return ok and self._graph is not None

    def run_fista_iterations(self, n_iters, rho, reg, num_row, num_feature, tol, w_tol=7e-5, fista_max_iter=5000):
        """up to n_iters ADMM iterations (stopping at the reference's stop test) in the library's native loop
        over the captured iteration graph; returns the rbl_run_stats, or None if the graph is not available"""
        if n_iters <= 0 or not self.graph_ready(w_tol, fista_max_iter):
            return None
        st = _cabi.RunStats()
        pyfloat = 1 if type(rho) is float else 0
        _cabi.check(self.lib.rbl_admm_run(self.h, ctypes.c_void_p(self._graph.raw_cuda_graph_exec()), self._stream(),
                                          ctypes.c_void_p(self.scal_host.data_ptr()),
                                          ctypes.c_void_p(self._out4_host.data_ptr()), int(n_iters), float(tol),
                                          float(reg), int(num_row), int(num_feature),
                                          int(self.active_dense_frac * self.n_local), float(rho), pyfloat,
                                          ctypes.byref(st)))
        self._graph_replays += st.iters
        self.Dw_valid, self._delta_valid, self._r_matches_w = True, False, False
        fs, ds, a = self.fista_stats, self.dual_stats, self.active_stats
        fs["calls"] += st.iters
        fs["iters"] += st.fista_iters
        fs["passes"] += st.fista_sweeps
        fs["last_passes"] = st.last_sweeps
        fs["d_passes"] += (st.iters - st.gathered) + st.dense_dual
        ds["sparse"] += st.sparse_dual
        ds["dense"] += st.dense_dual
        ds["nnz_last"] = st.nnz_last
        a["calls"] += st.iters
        a["rows"] += st.rows_read
        a["gathered"] += st.gathered
        return st

    def _capture_iteration(self, key):
        tol, max_iter = key
        dev = self.device
        self._ensure_scalars()
        torch.cuda.current_stream(dev).synchronize()
        stats = (dict(self.fista_stats), dict(self.dual_stats), dict(self.active_stats))
        g = torch.cuda.CUDAGraph()
        n0 = int(self.lib.rbl_launch_count())
        self._graph = None
        _cabi.check(self.lib.rbl_bind_scalars(self.h, self.scal.data_ptr()))
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_cap_stream"):
            self._cap_stream = torch.cuda.Stream(device=dev)
        try:
            # capture_begin / capture_end directly (torch.cuda.graph() would also run gc.collect and
            # empty_cache: tens of milliseconds in the middle of a solve)
            self._cap_stream.wait_stream(cur)
            with torch.cuda.stream(self._cap_stream):
                g.capture_begin()
                try:
                    self.scal.copy_(self.scal_host, non_blocking=True)
                    self.Dw_valid = True
                    self.z_step(1.0)                                   # by-value scalars are ignored while bound
                    self.w_step_fista(1.0, tol=tol, max_iter=max_iter)
                    self._dual_launch(1.0)
                finally:
                    g.capture_end()
            cur.wait_stream(self._cap_stream)
            self._graph, self._graph_key = g, key
            self._graph_launches, self._graph_replays = int(self.lib.rbl_launch_count()) - n0, 0
        except Exception as exc:  # noqa: BLE001 — capture unsupported here: stay on eager launches
            import warnings
            warnings.warn(f"rbl_b200: CUDA graph capture of the ADMM iteration failed ({exc!r}); "
                          "continuing with eager launches", RuntimeWarning, stacklevel=2)
            self._graph, self.graph_ok = None, False
        finally:
            _cabi.check(self.lib.rbl_bind_scalars(self.h, 0))
            # capture only records the launches: restore the bookkeeping it touched
            self.fista_stats, self.dual_stats, self.active_stats = stats
            self._active_pending = self._fista_info_pending = False
