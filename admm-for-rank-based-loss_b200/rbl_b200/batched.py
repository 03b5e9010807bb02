"""Batched mode (SURVEY.md K10, §8e "instance sharding"): many independent ADMM instances — a lambda
grid, several seeds / warm starts — that share ONE design matrix resident in HBM.

The reference has no counterpart (one `ADMMmethod` object per instance; the oracle for this mode is a
Python loop over it).  Two formulations:

* mode "gram" (default wherever the single-instance engine runs its w-step on G = D^T D): D, G and D^T are
  resident ONCE; every instance is a light child engine (own handle, state and scratch) whose whole ADMM
  iteration is a captured CUDA graph — active-row gradient gather, persistent FISTA on G, sparse dual pass —
  and the graphs of different instances are replayed concurrently on a pool of streams, so the latency-bound
  kernels of one instance (sort, PAV merge, FISTA barriers) overlap with the bandwidth-bound ones of another.
* mode "stream": the w-steps of all active instances advance together, every pass over D serving 8 instances
  at once (multi-RHS fused pass on the FP64 tensor-core path, csrc/batch_kernels.cu), each instance keeping
  its own device-resident FISTA state machine; z-steps run per instance on worker streams.

Either way each instance has its own rho schedule and stop test (ragged iteration counts), and across GPUs
instances are sharded with NO communication (`torch.distributed` only gathers the d x B result at the end).
"""
import os
import ctypes

import numpy as np
import torch

from . import _cabi
from . import spectra as _sp
from .engine import AdmmEngine, LOSS_IDS, _pow_table


class ZWorker:
    """Scratch + CUDA stream for the z-step (margins, radix sort, PAV prox, scatter) of one instance at a
    time.  The z-step never touches D, so several workers run the z-steps of different instances
    concurrently on their own streams while sharing nothing but read-only inputs."""

    def __init__(self, n, sigma, loss_id, clip, device):
        self.lib = _cabi.load()
        self.n, self.loss_id, self.clip, self.device = n, loss_id, clip, device
        h = ctypes.c_void_p()
        _cabi.check(self.lib.rbl_create(ctypes.byref(h), device.index or 0, n, n, 0, 2, 2))
        self.h = h
        self.stream = torch.cuda.Stream(device=device)
        f64 = torch.float64
        self.m = torch.empty(n, dtype=f64, device=device)
        self.m_sorted = torch.empty(n, dtype=f64, device=device)
        self.z_sorted = torch.empty(n, dtype=f64, device=device)
        self.perm = torch.empty(n, dtype=torch.int32, device=device)
        with torch.cuda.stream(self.stream):
            _cabi.check(self.lib.rbl_set_spectrum(self.h, sigma.data_ptr(), ctypes.c_void_p(self.stream.cuda_stream)))
        self.stream.synchronize()

    def z_step(self, Dw, lam, rho, z_out, b_out):
        lib, h, s = self.lib, self.h, ctypes.c_void_p(self.stream.cuda_stream)
        _cabi.check(lib.rbl_margins(h, Dw.data_ptr(), lam.data_ptr(), rho, self.m.data_ptr(), s))
        _cabi.check(lib.rbl_sort_margins(h, self.m.data_ptr(), self.m_sorted.data_ptr(), self.perm.data_ptr(), s))
        _cabi.check(lib.rbl_pav_prox(h, self.loss_id, self.m_sorted.data_ptr(), rho, self.z_sorted.data_ptr(), s))
        _cabi.check(lib.rbl_scatter_z(h, self.z_sorted.data_ptr(), self.perm.data_ptr(), 0 if self.clip is None else 1,
                                      0.0 if self.clip is None else float(self.clip), lam.data_ptr(), rho,
                                      z_out.data_ptr(), b_out.data_ptr(), s))

    def close(self):
        if self.h is not None and self.h.value:
            self.lib.rbl_destroy(self.h)
            self.h = None


def instance_shard(B, world, rank):
    """contiguous block of instances for this rank (sizes differ by at most one)"""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class BatchedADMM:
    """Solve  min_w sum_i sigma_i loss_[i](D w) + l1_j/2 |w|_1  for every l1_j in `l1_regs`
    (optionally with per-instance warm starts `w0s`), all sharing X, y, the spectrum and the loss.

    Iterates are those of `ADMMmethod` run separately per instance (same state initialisation
    src/optim/algorithms.py:30-52, same iteration :119-164, FISTA w-step :190-202).
    """

    def __init__(self, X, y, weight_function="erm", loss="binary_cross_entropy", l1_regs=None, B_clip=None, args=None,
                 w0s=None, max_iter=200, tol=1e-4, device=None, group=None, shard=True, z_streams=8, mode=None):
        if l1_regs is None or len(l1_regs) == 0:
            raise ValueError("l1_regs: one l1 regulariser per instance is required (batched mode is the l1/FISTA path)")
        if B_clip is not None and weight_function != "ehrm":
            raise ValueError(f"Unrecognized weight_function '{weight_function}'! Options: ['ehrm']")
        X = np.asarray(X) if not torch.is_tensor(X) else X
        self.n, self.d = int(X.shape[0]), int(X.shape[1])
        self.B_total = len(l1_regs)
        dist = torch.distributed
        self.world = dist.get_world_size(group) if (shard and dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.group = group
        self.i_lo, self.i_hi = instance_shard(self.B_total, self.world, self.rank)
        self.regs = np.asarray(l1_regs, dtype=np.float64)[self.i_lo:self.i_hi]
        self.B = len(self.regs)
        wf = _sp.get_weights(weight_function, args)
        sig_a, sig_b = (wf[0](self.n), wf[1](self.n)) if isinstance(wf, tuple) else (wf(self.n), None)
        self.sigma_a = sig_a
        self.loss = loss
        if weight_function == "ehrm":
            # the reference's per-z-step choice between its two clipped candidates (PAV_cpt.py:203-226) needs the host
            # between the sort and the prox of every instance: served by the per-instance engines of the gram mode
            # (their iterations then run eagerly, not as replayed graphs); the multi-RHS stream mode has no such hook
            if B_clip is None:
                raise TypeError("weight_function 'ehrm' needs B_clip (PAV_cpt.py:207 compares the prox with it)")
            if (mode or os.environ.get("RBL_BATCH_MODE", "auto")).lower() == "stream":
                raise ValueError("batched EHRM runs in mode='gram' (the candidate choice is made per instance)")
            mode = "gram"
            self.eng = AdmmEngine(X, y, loss, None, ehrm=(sig_a, sig_b, float(B_clip)), device=device)
        else:
            self.eng = AdmmEngine(X, y, loss, sig_a, device=device)
        e = self.eng
        self.tol, self.max_iter = tol, max_iter
        self.rho0 = 1e-4 if weight_function == "ehrm" else (2e-7 if weight_function in ("aorr", "aorr_dc") else 1e-5)
        B, n, d, dev = max(self.B, 1), self.n, self.d, e.device
        f64 = torch.float64
        self.W = torch.zeros((B, d), dtype=f64, device=dev)
        self.W_prev = torch.zeros((B, d), dtype=f64, device=dev)
        self.Z = torch.zeros((B, n), dtype=f64, device=dev)
        self.LAM = torch.zeros((B, n), dtype=f64, device=dev)
        self.DW = torch.zeros((B, n), dtype=f64, device=dev)
        self.Bv = torch.zeros((B, n), dtype=f64, device=dev)
        self.R = torch.zeros((B, n), dtype=f64, device=dev)
        self.out4 = torch.zeros((B, 4), dtype=f64, device=dev)
        self.out4_host = torch.zeros((B, 4), dtype=f64).pin_memory()
        for j in range(self.B):
            reg = float(self.regs[j])
            lam0 = 0.1 * reg / n
            self.Z[j].fill_(lam0)
            self.LAM[j].fill_(lam0)
            if w0s is not None:
                self.W[j].copy_(e.vec(np.asarray(w0s)[self.i_lo + j]))
            else:
                self.W[j].fill_(0.001 * reg / d / n)
        self.rho = [self.rho0] * self.B          # python floats first, np.float64 after the first update
        self.slot_of = list(range(self.B))       # slot -> local instance id (finished instances move to the end)
        self.n_active = self.B
        self.iters = np.zeros(self.B, dtype=np.int64)
        self.converged = np.zeros(self.B, dtype=bool)
        self.primal = np.full(self.B, np.inf)
        self.dual = np.full(self.B, np.inf)
        self.fista_passes = 0   # passes over D (each serves up to 8 instances)
        mode = (mode or os.environ.get("RBL_BATCH_MODE", "auto")).lower()
        if mode not in ("auto", "gram", "stream"):
            raise ValueError(f"mode must be auto, gram or stream (got {mode!r})")
        self.mode = e.w_mode if mode == "auto" else mode
        if self.mode == "gram" and e.w_mode != "gram":
            raise ValueError("batched mode 'gram' needs the engine in Gram mode (d <= 4096, n >= 2d)")
        if self.mode == "gram":
            self._init_gram(w0s, z_streams)
            return
        if self.B:
            _cabi.check(e.lib.rbl_batch_create(e.h, self.B))
            tab = _pow_table(np.float32(2.5))
            _cabi.check(e.lib.rbl_fista_config(e.h, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
            for j in range(self.B):  # D w0 per instance
                e.matvec(self.W[j], self.DW[j])
            torch.cuda.current_stream(e.device).synchronize()
            self.zworkers = [ZWorker(n, e.sigma, e.loss_id, e.clip, e.device) for _ in range(min(z_streams, self.B))]

    # ---- Gram formulation: one child engine per instance, graphs replayed concurrently ----------------
    def _init_gram(self, w0s, n_streams):
        e, n, d = self.eng, self.n, self.d
        # the big per-slot matrices of the stream formulation are not needed
        self.Z = self.LAM = self.DW = self.Bv = self.R = None
        self.inst = []
        # RBL_BATCH_PASS_GRID: persistent CTAs of every instance's D-reading kernels (default: one per SM).  Giving
        # each instance a quarter of the SMs so that gathers overlap other instances' latency-bound kernels was
        # measured and does not help (32 instances x 100k x 1000: 7.99 ms per batched step at 148 CTAs, 8.15 at 37,
        # 9.0 at 24 / 16): the per-instance graphs do not overlap for another reason
        share = int(os.environ.get("RBL_BATCH_PASS_GRID", "0")) or int(e.info["num_sms"])
        self.pass_grid_per_instance = share
        for j in range(self.B):
            c = AdmmEngine.child(e)
            _cabi.check(c.lib.rbl_set_pass_grid(c.h, share))
            reg = float(self.regs[j])
            lam0 = 0.1 * reg / n
            w_init = (np.asarray(w0s)[self.i_lo + j] if w0s is not None else np.full(d, 0.001 * reg / d / n))
            c.set_state(w=w_init, z=np.full(n, lam0), lam=np.full(n, lam0))
            self.inst.append(c)
        self.active = list(range(self.B))
        self.streams = [torch.cuda.Stream(device=e.device) for _ in range(max(1, min(n_streams, self.B)))]
        torch.cuda.current_stream(e.device).synchronize()

    def _step_gram(self):
        e = self.eng
        if not self.active:
            return 0
        main = torch.cuda.current_stream(e.device)
        ready = torch.cuda.Event()
        ready.record(main)
        for k, j in enumerate(self.active):
            st = self.streams[k % len(self.streams)]
            if k < len(self.streams):
                st.wait_event(ready)
            reg = float(self.regs[j])
            alpha = reg / (2 * self.rho[j] * self.n)   # algorithms.py:192-193,200
            with torch.cuda.stream(st):
                self.inst[j].iteration_fista(self.rho[j], alpha * self.n, sync=False)
        for st in self.streams:
            st.synchronize()
        still = []
        for j in self.active:
            pf, df = self.inst[j].finish_iteration()
            self.primal[j], self.dual[j] = pf, df
            self.iters[j] += 1
            if pf < self.tol and df < self.tol:
                self.converged[j] = True
            else:
                self.rho[j] = np.min((self.rho[j] * (1.02 if pf > 1e-2 else 1.07), 217 * self.d))
                still.append(j)
        self.active = still
        self.n_active = len(still)
        return self.n_active

    @property
    def last_fista_info(self):
        """{instance: (FISTA iterations, 1 + line-search trials, final L)} of the last step"""
        if self.mode != "gram":
            return self._last_fista_info
        out = {}
        hi, hd = (ctypes.c_int32 * 8)(), (ctypes.c_double * 4)()
        for j, c in enumerate(self.inst):
            _cabi.check(c.lib.rbl_fista_poll(c.h, c._stream(), hi, hd))
            out[j] = (int(hi[1]), 1 + int(hi[3]), float(hd[1]))
        return out

    # ------------------------------------------------------------------------------------------
    def _swap_slots(self, a, b):
        if a == b:
            return
        for T in (self.W, self.W_prev, self.Z, self.LAM, self.DW, self.Bv, self.R):
            tmp = T[a].clone()
            T[a].copy_(T[b])
            T[b].copy_(tmp)
        self.slot_of[a], self.slot_of[b] = self.slot_of[b], self.slot_of[a]
        self.rho[a], self.rho[b] = self.rho[b], self.rho[a]

    def step(self):
        """one ADMM iteration of every active instance; returns the number still active"""
        if self.mode == "gram":
            return self._step_gram()
        e, lib = self.eng, self.eng.lib
        Ba = self.n_active
        if Ba == 0:
            return 0
        s = e._stream
        # ---- z-steps, per instance (algorithms.py:88-106), round-robin over the z-workers' streams
        main = torch.cuda.current_stream(e.device)
        ready = torch.cuda.Event()
        ready.record(main)
        for zw in self.zworkers:
            zw.stream.wait_event(ready)
        for a in range(Ba):
            zw = self.zworkers[a % len(self.zworkers)]
            zw.z_step(self.DW[a], self.LAM[a], float(self.rho[a]), self.Z[a], self.Bv[a])
        for zw in self.zworkers:
            done_ev = torch.cuda.Event()
            done_ev.record(zw.stream)
            main.wait_event(done_ev)
        # ---- w-steps, all active instances together (algorithms.py:190-202, fast_lasso.py:22-69)
        self.W_prev[:Ba].copy_(self.W[:Ba])
        lams = (ctypes.c_double * Ba)()
        flags = (ctypes.c_int32 * Ba)()
        for a in range(Ba):
            reg = float(self.regs[self.slot_of[a]])
            alpha = reg / (2 * self.rho[a] * self.n)
            lam = alpha * self.n
            lams[a] = float(lam)
            flags[a] = 1 if type(lam) is float else 0  # NEP 50: python-float lam -> float32 threshold (iteration 0)
        _cabi.check(lib.rbl_fista_batch_begin(e.h, Ba, self.W_prev.data_ptr(), lams, flags, 17.0, 7e-5, 5000, s()))
        done = (ctypes.c_int32 * Ba)()
        its = (ctypes.c_int32 * Ba)()
        passes = (ctypes.c_int32 * Ba)()
        Ls = (ctypes.c_double * Ba)()
        batch = max(4, getattr(self, "_last_steps", 0) - 1)
        steps = 0
        while True:
            _cabi.check(lib.rbl_fista_batch_steps(e.h, Ba, e.D.data_ptr(), self.Bv.data_ptr(), batch, s()))
            steps += batch
            _cabi.check(lib.rbl_fista_batch_poll(e.h, Ba, s(), done, its, passes, Ls))
            if all(done[a] for a in range(Ba)):
                break
            batch = 4
        self._last_steps = max(passes[a] for a in range(Ba))
        self.last_fista_iters = {self.slot_of[a]: int(its[a]) for a in range(Ba)}
        self._last_fista_info = {self.slot_of[a]: (int(its[a]), int(passes[a]), float(Ls[a])) for a in range(Ba)}
        self.fista_passes += self._last_steps * ((Ba + 7) // 8)
        _cabi.check(lib.rbl_fista_batch_result(e.h, Ba, self.W.data_ptr(), self.R.data_ptr(), s()))
        # ---- dual updates + residuals, per instance (algorithms.py:132-136)
        for a in range(Ba):
            _cabi.check(lib.rbl_dual_update(e.h, self.Z[a].data_ptr(), self.DW[a].data_ptr(), self.Bv[a].data_ptr(),
                                            self.R[a].data_ptr(), 1, self.LAM[a].data_ptr(), float(self.rho[a]),
                                            self.W[a].data_ptr(), self.W_prev[a].data_ptr(),
                                            self.out4[a].data_ptr(), s()))
        self.out4_host.copy_(self.out4, non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        o = self.out4_host.numpy()
        finished = []
        for a in range(Ba):
            j = self.slot_of[a]
            pf, df = float(np.sqrt(o[a, 0])), float(np.sqrt(o[a, 1]))
            self.primal[j], self.dual[j] = pf, df
            self.iters[j] += 1
            if pf < self.tol and df < self.tol:
                self.converged[j] = True
                finished.append(a)
            else:
                mult = 1.02 if pf > 1e-2 else 1.07
                self.rho[a] = np.min((self.rho[a] * mult, 217 * self.d))
        for a in sorted(finished, reverse=True):  # retire: swap into the tail of the active range
            self._swap_slots(a, self.n_active - 1)
            self.n_active -= 1
        return self.n_active

    def main_loop(self, verbose=False):
        for it in range(self.max_iter):
            left = self.step()
            if verbose and it % 10 == 0:
                print("iter_num=", it, "active instances:", left)
            if left == 0:
                break
        return self.result()

    def close(self):
        for zw in getattr(self, "zworkers", []):
            zw.close()
        for c in getattr(self, "inst", []):
            c.close()
        self.eng.close()

    def result(self, gather=True):
        """d x B matrix of solutions, columns in the order of `l1_regs` (all ranks' instances if gathered)"""
        W_local = np.zeros((self.d, self.B))
        if self.mode == "gram":
            for j, c in enumerate(self.inst):
                W_local[:, j] = c.w.cpu().numpy()
        else:
            Wh = self.W.cpu().numpy()
            for slot, j in enumerate(self.slot_of):
                W_local[:, j] = Wh[slot]
        if self.world == 1 or not gather:
            return W_local
        out = [None] * self.world
        torch.distributed.all_gather_object(out, W_local, group=self.group)
        return np.concatenate(out, axis=1)

    def state(self, j):
        """(w, z, lambda, rho) of local instance j as numpy arrays (for inspection / lockstep tests)"""
        if self.mode == "gram":
            c = self.inst[j]
            return (c.w.cpu().numpy(), c.z.cpu().numpy(), c.lam.cpu().numpy(), self.rho[j])
        slot = self.slot_of.index(j)
        return (self.W[slot].cpu().numpy(), self.Z[slot].cpu().numpy(), self.LAM[slot].cpu().numpy(), self.rho[slot])

    def objective(self, j):
        """rank-weighted objective of local instance j (objective.py:71-87)"""
        e = self.eng
        if not hasattr(self, "_sig_dev"):
            self._sig_dev = e.vec(self.sigma_a)
        wj = self.inst[j].w if self.mode == "gram" else self.W[self.slot_of.index(j)].contiguous()
        risk, w2, w1 = e.objective_terms(wj, self._sig_dev, self.loss)
        return risk + 0.5 * float(self.regs[j]) * w1
