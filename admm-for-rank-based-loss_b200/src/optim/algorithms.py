"""Drop-in for the reference's src/optim/algorithms.py (ADMM entry point used by run_demo.py and the
run_*.py drivers): same class names, constructor/method signatures, attributes and error behaviour —
the numerical work runs on the B200 through librbl_b200.so (no CPU fallback).

Per iteration (reference :119-164):
    z-step  margins -> stable radix sort -> PAV/isotonic prox -> scatter        (device, :88-106)
    w-step  l1: FISTA as a device-resident state machine, one fused pass over D per trial (:190-202)
            l2: scipy L-BFGS-B on the host driving a fused device f/g pass        (:109-116, w_LBFGS.py)
    dual    lambda += rho (z - D w), residual norms, stop test, rho schedule     (device + host scalars)

EHRM: the reference's all-or-nothing choice between its two clipped candidates (PAV_cpt.py:222-226) is made on
the device from the same two sums at every z-step.

Documented deviations from the shipped reference: exact PAV + machine-precision
Newton in the z-step (the reference's sweep PAV / loose Newton converge to the same unique prox);
closed-form hinge prox; stable sort; FISTA in float64 by default (`fista_dtype` is accepted for API
compatibility; float32 is the shipped default there and is chaotic at its own tolerance).  The
tiny-problem sklearn-Lasso branch (:194-197) runs scikit-learn's coordinate descent on the device.
"""
import time

import numpy as np
import torch

from rbl_b200.engine import AdmmEngine
from src.optim.objective import rankbasedObjective


class Optimizer:
    def __init__(self, X, y, weight_function="erm", loss="binary_cross_entropy", l2_reg=None, l1_reg=None,
                 B=None, n_class=None, args=None, w0=None, max_iter=200, tol=1e-4, _shard=None, _storage=None,
                 _share=None):
        # _share (extension): another solver (or its engine) built on the SAME X, y — this one then borrows its
        # device-resident D, G = D^T D and D^T (no second upload, no second SYRK) and only owns its state, spectrum
        # and scratch: the ratio / spectrum / regulariser sweeps of run_AoRR_ratio.py:32-46 and run_SRM.py on one
        # design matrix.  X and y may be None then.  The first solver must outlive the ones sharing with it.
        # _shard (extension, not in the reference): dict(row_lo=, n_global=[, group=]) when X, y are this
        # rank's contiguous rows of a row-sharded problem (one process per GPU, torch.distributed/NCCL)
        # _storage (extension): "fp32" keeps D in float32 in HBM (optional mode: half the bytes per pass, fp64
        # arithmetic throughout; iterates within ~1e-6 of the fp64 run) — default "fp64" (env RBL_STORAGE)
        _t0 = time.perf_counter()
        _shard = dict(_shard or {})
        if _share is not None:
            parent = getattr(_share, "engine", _share)
            X, y = _ShapeOnly((parent.n_local, parent.d)), None
            _shard = {"_share": parent}
            if parent.world > 1:
                _shard["n_global_hint"] = parent.n_global
        elif not torch.is_tensor(X):  # a (device) tensor is taken as it is: no copy back to the host
            X = np.asarray(X)
        if _share is None:
            y = y.detach().cpu().numpy() if torch.is_tensor(y) else np.asarray(y)
        if _storage is not None:
            _shard["storage"] = _storage
        self.num_row = int(_shard.pop("n_global_hint", _shard.get("n_global", X.shape[0])))
        self.num_feature = X.shape[1]
        # regularization (:30) — raises TypeError below when both are None, like the reference (:32)
        self.reg = l1_reg or l2_reg
        self.loss = loss
        # objective first: it validates weight_function / loss / args exactly like the reference (:22)
        self.objective = _LazyObjective(X, y, weight_function, loss, l2_reg, l1_reg, B, n_class, args,
                                        self.num_row)
        self.sigma_a = self.objective.alphas.numpy().reshape(-1)
        self.sigma_b = self.objective.betas.numpy().reshape(-1)
        lam0 = 0.1 * self.reg / self.num_row  # :32,34
        if w0 is not None:
            w_init = np.asarray(w0, dtype=np.float64).reshape(-1, 1)
        else:
            w_init = 0.001 * self.reg / self.num_feature / self.num_row * np.ones(shape=(self.num_feature, 1))
        self.tol = tol
        self.max_iter = max_iter
        if weight_function == 'ehrm':
            self.rho = 0.0001
        elif weight_function == 'aorr' or weight_function == 'aorr_dc':
            self.rho = 2e-7
        else:
            self.rho = 1e-5
        if l1_reg is not None:
            self.w_flag = 1
        elif l2_reg is not None:
            self.w_flag = 2
        else:
            raise ValueError("More arguments: l1_reg or l2_reg not l1_reg and l2_reg!")
        self.B = B
        if B is not None and weight_function != "ehrm":
            raise ValueError(
                f"Unrecognized weight_function '{weight_function}'! Options: ['ehrm']"
            )
        self.w_tol = 7e-5
        self.z_maxiter = self.num_row
        self.store = False
        self.weight_function = weight_function

        # EHRM (PAV_cpt.py:203-293): every z-step compares the reference's two scalar sums on the device and runs
        # the winner, min(B, isotonic prox with sigma = alphas) or max(B, isotonic prox with sigma = betas)
        eX, ey = (None, None) if _share is not None else (X, y)
        if weight_function == 'ehrm':
            if B is None:
                raise TypeError("weight_function 'ehrm' needs B (PAV_cpt.py:207 compares the prox with it)")
            self.engine = AdmmEngine(eX, ey, loss, None, ehrm=(self.sigma_a, self.sigma_b, float(B)), **_shard)
        else:
            self.engine = AdmmEngine(eX, ey, loss, self.sigma_a, **_shard)
        self.objective._attach(self.engine)
        nl = self.engine.n_local
        self.engine.set_state(w=w_init, z=np.full(nl, lam0), lam=np.full(nl, lam0))
        self._w = w_init.astype(np.float64)
        if hasattr(self.engine, "build_times"):
            self.engine.build_times["constructor_total_s"] = time.perf_counter() - _t0
        self.fista_max_iter = 5000
        self.last_info = {}

    # ---- state views (numpy, like the reference's attributes) -----------------------------------
    @property
    def w(self):
        return self._w

    @w.setter
    def w(self, value):
        self._w = np.asarray(value, dtype=np.float64).reshape(-1, 1)
        self.engine.set_state(w=self._w)

    @property
    def z(self):
        return self.engine.z.cpu().numpy().reshape(-1, 1)

    @z.setter
    def z(self, value):
        self.engine.set_state(z=value)

    @property
    def lagrangian(self):
        return self.engine.lam.cpu().numpy().reshape(-1, 1)

    @lagrangian.setter
    def lagrangian(self, value):
        self.engine.set_state(lam=value)

    @property
    def D(self):
        return self.engine.D[:, : self.num_feature].cpu().numpy().astype(np.float64, copy=False)

    @property
    def DTD(self):
        D = self.D
        return D.T @ D

    def start_store(self, X, y, weight_function="erm", loss="binary_cross_entropy",
                    B=None, l2_reg=None, l1_reg=None, n_class=None, args=None):
        # X, y both are test set.
        Xt = X if torch.is_tensor(X) else torch.from_numpy(np.asarray(X))
        yt = y if torch.is_tensor(y) else torch.from_numpy(np.asarray(y))
        self.test_objective = rankbasedObjective(Xt, yt,
                                                 weight_function, loss, l2_reg, l1_reg, B, n_class, args)
        self.w_time = [0]
        self.z_time = [0]
        self.train_losses = [self.objective.get_arrogate_loss(torch.from_numpy(self.w).double())]
        self.test_losses = [self.test_objective.get_arrogate_loss(torch.from_numpy(self.w).double())]
        self.time_array = [0]
        self.store = True

    def z_subproblem(self):
        self.engine.z_step(self.rho)
        return self.z

    def _rebuild_b(self):
        """b = z + lagrangian / rho from the CURRENT state, as the reference forms it at the top of every w-step
        (:111, w_LBFGS.py:34, :191) — the engine's b is only the by-product of the last z-step's scatter"""
        e = self.engine
        e._delta_valid = False  # not the z-step's z - m any more: the gradient pass reads every row
        e._pre_done = False
        e.b.copy_(e.z + e.lam / float(self.rho))

    def w_subproblem(self):
        """reference :109-116 — standalone l2 w-step on the current (z, lagrangian, rho); returns numpy d x 1"""
        if self.w_flag == 2:
            self._rebuild_b()
            self.last_info = self.engine.w_step_lbfgs(self.rho, self.reg)
        else:
            raise ValueError("w_flag can only be 0, 1 or 2.")
        return self.engine.w.cpu().numpy().reshape(-1, 1)

    def _sync_timer(self):
        if self.store:
            torch.cuda.synchronize(self.engine.device)
        return time.time()

    def main_loop(self, i, t_start, verbose):
        if not self.store and self._whole_iteration_on_device():
            # z-step, FISTA w-step and dual step as one device-side sequence (a CUDA graph after warm-up)
            alpha = self.reg / (2 * self.rho * self.num_row)
            primal_feasibility, dual_feasibility = self.engine.iteration_fista(
                self.rho, alpha * self.num_row, tol=self.w_tol, max_iter=self.fista_max_iter)
        else:
            t1 = self._sync_timer()
            if not self.store and self._w_step_is_lbfgs():
                # z-step and the warm-start pass of the L-BFGS w-step as one device-side sequence (a CUDA graph
                # after warm-up); with start_store the two steps stay apart so that z_time / w_time keep their
                # meaning
                self.engine.z_and_grad(self.rho)
            else:
                self.engine.z_step(self.rho)
            t2 = self._sync_timer()
            if self.store:
                self.z_time.append(t2 - t1 + self.z_time[i])

            self._w_subproblem_device()
            if self.store:
                self.w_time.append(self._sync_timer() - t2 + self.w_time[i])

            # Lagrange multiplier update + stopping criterion (:132-136), fused on the device
            primal_feasibility, dual_feasibility = self.engine.dual_step(self.rho)
        self._w = self.engine.w_host.numpy().reshape(-1, 1).copy()
        self.primal_feasibility, self.dual_feasibility = primal_feasibility, dual_feasibility
        if primal_feasibility < self.tol and dual_feasibility < self.tol:
            print('algorithm converges within tolerance')
            print('iter_num=', i, 'primal_feasibility: ', primal_feasibility, 'dual_feasibility: ', dual_feasibility)
            print('loss=', self.objective.get_arrogate_loss(torch.from_numpy(self.w).double()))
            return True
        if verbose:
            if i % 10 == 0:
                print('iter_num=', i, 'primal_feasibility: ', primal_feasibility, 'dual_feasibility: ', dual_feasibility)
                print('loss=', self.objective.get_arrogate_loss(torch.from_numpy(self.w).double()))

        # ALM penalty update (:147-157; the 'ehrm'/'aorr'/'aorr_dc' branches compare self.loss — a loss
        # name — against weight-function names and never fire; reproduced as is)
        if self.loss == 'ehrm' or self.loss == 'aorr':
            self.rho = np.min((self.rho * 1.2, 17 * self.num_feature))
        elif self.loss == 'aorr_dc':
            if i >= 7 and i % 3 == 0:
                self.rho = np.min((self.rho * 5, 17 * self.num_feature))
        else:
            if primal_feasibility > 1e-2:
                self.rho = np.min((self.rho * 1.02, 217 * self.num_feature))
            else:
                self.rho = np.min((self.rho * 1.07, 217 * self.num_feature))

        if self.store:
            self.train_losses.append(self.objective.get_arrogate_loss(torch.from_numpy(self.w).double()))
            self.test_losses.append(self.test_objective.get_arrogate_loss(torch.from_numpy(self.w).double()))
            self.time_array.append(time.time() - t_start)

        return False

    def _whole_iteration_on_device(self):
        return False

    def _w_step_is_lbfgs(self):
        return self.w_flag == 2

    def _w_subproblem_device(self):
        """w-step leaving the result on the device (the numpy view is refreshed by the dual step)."""
        if self.w_flag == 2:
            self.last_info = self.engine.w_step_lbfgs(self.rho, self.reg)
        else:
            raise ValueError("w_flag can only be 0, 1 or 2.")

    def final_res(self):
        if self.store:
            return self.w, self.time_array, self.train_losses, self.test_losses
        else:
            raise ValueError("Data was not saved.")


class _LazyObjective(rankbasedObjective):
    """rankbasedObjective that shares the ADMM engine's device-resident D instead of building its own."""

    def __init__(self, X, y, weight_function, loss, l2_reg, l1_reg, B, n_class, args, n_global):
        super().__init__(X, y, weight_function, loss, l2_reg, l1_reg, B, n_class, args, _problem=_Deferred(),
                         _n=n_global)

    def _attach(self, engine):
        self.problem = engine
        self._alphas_dev = engine.vec(self.alphas)

    def get_arrogate_loss(self, w, include_reg=True):
        """objective.py:71-87.  Right after a dual step the engine holds D w for exactly this w (the per-iteration
        train loss of start_store, algorithms.py:160): the margins are reused and no pass over D is made."""
        eng = self.problem
        wt = w if torch.is_tensor(w) else torch.from_numpy(np.ascontiguousarray(w))
        if (getattr(eng, "Dw_valid", False) and not wt.is_cuda and wt.numel() == eng.d
                and torch.equal(wt.reshape(-1).to(torch.float64), eng.w_host.reshape(-1))):
            risk, w2, w1 = eng.objective_terms(eng.w, self._alphas_dev, self.loss_name, u_local=eng.Dw)
            if self.l2_reg and include_reg:
                risk += 0.5 * self.l2_reg * w2
            if self.l1_reg and include_reg:
                risk += 0.5 * self.l1_reg * w1
            return risk
        return super().get_arrogate_loss(w, include_reg)


class _Deferred:
    def vec(self, a):
        return None


class _ShapeOnly:
    """stands in for X when the design matrix is borrowed from another solver (_share)"""

    def __init__(self, shape):
        self.shape = shape


class ADMMmethod(Optimizer):
    def __init__(self, X, y, weight_function="erm", loss="binary_cross_entropy",
                 l2_reg=None, l1_reg=None, B=None, n_class=None, args=None, w0=None, max_iter=200, tol=1e-4,
                 _shard=None, _storage=None, _share=None):
        super(ADMMmethod, self).__init__(X, y, weight_function, loss, l2_reg, l1_reg, B, n_class,
                                         args, w0, max_iter, tol, _shard, _storage, _share)

    def start_store(self, X, y, weight_function="erm", loss="binary_cross_entropy",
                    B=None, l2_reg=None, l1_reg=None, n_class=None, args=None):
        super(ADMMmethod, self).start_store(X, y, weight_function, loss, B,
                                            l2_reg, l1_reg, n_class, args)

    def _z_subproblem(self):
        return super(ADMMmethod, self).z_subproblem()

    def _small_lasso(self):
        return self.num_row <= 500 and self.num_feature <= 60  # :194

    def _whole_iteration_on_device(self):
        return self.w_flag == 1 and self.engine.w_mode == "gram" and not self._small_lasso()

    def _w_subproblem_device(self):
        if self.w_flag == 1:
            # const_y = z + lambda/rho was written by the scatter kernel; lam = alpha*n = reg/(2 rho) (:192-193,200)
            alpha = self.reg / (2 * self.rho * self.num_row)
            if self._small_lasso():
                # :194-197 — sklearn.linear_model.Lasso(alpha, tol=1e-8, fit_intercept=False, max_iter=50000):
                # its coordinate descent, on the device
                self.last_info = self.engine.w_step_lasso_cd(alpha * self.num_row, tol=1e-8, max_iter=50000)
                return
            self.last_info = self.engine.w_step_fista(alpha * self.num_row, tol=self.w_tol,
                                                      max_iter=self.fista_max_iter)
        elif self.w_flag == 0 or self.w_flag == 2:
            super(ADMMmethod, self)._w_subproblem_device()
        else:
            raise ValueError("w_flag can only be 0, 1 or 2.")

    def _w_subproblem(self):
        """reference :190-207 — standalone w-step on the current (z, lambda, rho); returns numpy d x 1"""
        self._rebuild_b()
        self._w_subproblem_device()
        return self.engine.w.cpu().numpy().reshape(-1, 1)

    def advance(self, i0, n_iters, verbose=False, t_start=0.0):
        """Iterations i0 .. i0 + n_iters - 1 of the reference loop (:209-216); returns (next i, converged).
        Extension used by main_loop and bench.py: once the engine has captured the iteration as a CUDA graph the
        iterations between two verbose prints run in the library's native loop (rbl_admm_run) — same host logic
        as Optimizer.main_loop (stop test, rho schedule), no interpreter in between."""
        i, end = i0, i0 + n_iters
        while i < end:
            st = None
            if not self.store and self._whole_iteration_on_device():
                # run up to and including the next iteration that prints (i % 10 == 0), natively
                stop = end if not verbose else min(end, i + 1 if i % 10 == 0 else (i // 10 + 1) * 10 + 1)
                st = self.engine.run_fista_iterations(stop - i, self.rho, self.reg, self.num_row, self.num_feature,
                                                      self.tol, self.w_tol, self.fista_max_iter)
            elif (not self.store and self.w_flag == 2 and self.engine.w_mode == "gram"
                  and self.engine.ehrm is None):
                # l2 problems: graph (z-step + gradient pass) -> the library's L-BFGS-B -> graph (dual step), natively
                stop = end if not verbose else min(end, i + 1 if i % 10 == 0 else (i // 10 + 1) * 10 + 1)
                st = self.engine.run_lbfgs_iterations(stop - i, self.rho, self.reg, self.num_feature, self.tol)
                if st is not None:
                    self.last_info = {"nit": None, "nfev": int(st.last_sweeps), "solver": "librbl_b200 (native loop)"}
            if st is None:
                if super(ADMMmethod, self).main_loop(i, t_start, verbose):
                    return i + 1, True
                i += 1
                continue
            i += st.iters
            last = i - 1
            self._w = self.engine.w_host.numpy().reshape(-1, 1).copy()
            self.primal_feasibility, self.dual_feasibility = st.primal, st.dual
            if not st.rho_is_pyfloat:
                self.rho = np.float64(st.rho)
            if st.converged:
                print('algorithm converges within tolerance')
                print('iter_num=', last, 'primal_feasibility: ', st.primal, 'dual_feasibility: ', st.dual)
                print('loss=', self.objective.get_arrogate_loss(torch.from_numpy(self.w).double()))
                return i, True
            if verbose and last % 10 == 0:
                print('iter_num=', last, 'primal_feasibility: ', st.primal, 'dual_feasibility: ', st.dual)
                print('loss=', self.objective.get_arrogate_loss(torch.from_numpy(self.w).double()))
        return i, False

    def main_loop(self, verbose=True):
        t_start = time.time()
        self.advance(0, self.max_iter, verbose, t_start)
        return self.w

    def final_res(self):
        return super(ADMMmethod, self).final_res()


class smoothADMMmethod(Optimizer):
    """reference :223-263 — ADMM with the l1 term replaced by its Huber-type smoothing
    reg/2 * (w^2/(2t) if |w| <= t else |w| - t/2) (w_LBFGS.py:11-28), minimised by L-BFGS-B; t follows the
    reference's schedule (:255) and the final iterate is soft-thresholded at t (:258)."""

    def __init__(self, X, y, weight_function="erm", loss="binary_cross_entropy",
                 B=None, l2_reg=None, l1_reg=None, n_class=None, args=None, w0=None, t=1, max_iter=200, tol=1e-4,
                 _shard=None, _storage=None):
        super(smoothADMMmethod, self).__init__(X, y, weight_function, loss, l2_reg, l1_reg, B, n_class,
                                               args, w0, max_iter, tol, _shard, _storage)
        self.t = t

    def start_store(self, X, y, weight_function="erm", loss="binary_cross_entropy",
                    B=None, l2_reg=None, l1_reg=None, n_class=None, args=None):
        super(smoothADMMmethod, self).start_store(X, y, weight_function, loss, B,
                                                  l2_reg, l1_reg, n_class, args)

    def _z_subproblem(self):
        return super(smoothADMMmethod, self).z_subproblem()

    def _w_step_is_lbfgs(self):
        return self.w_flag in (1, 2)  # the Huber-smoothed l1 problem goes through L-BFGS-B as well (:247-251)

    def _w_subproblem(self):
        """reference :237-251 — standalone w-step on the current (z, lagrangian, rho, t); returns numpy d x 1"""
        self._rebuild_b()
        self._w_subproblem_device()
        return self.engine.w.cpu().numpy().reshape(-1, 1)

    def _w_subproblem_device(self):
        if self.w_flag == 1:
            # wl1_fun_smooth / wl1_fun_smooth_gradient (w_LBFGS.py:11-28) minimised by L-BFGS-B (:54-62)
            self.last_info = self.engine.w_step_lbfgs(self.rho, self.reg, huber_t=float(self.t))
        elif self.w_flag == 0 or self.w_flag == 2:
            super(smoothADMMmethod, self)._w_subproblem_device()
        else:
            raise ValueError("w_flag can only be 0, 1 or 2.")

    def main_loop(self, verbose=True):
        t_start = time.time()

        for i in range(self.max_iter):
            if super(smoothADMMmethod, self).main_loop(i, t_start, verbose):
                break
            if i >= 17:
                self.t = max(self.t * 0.9, 1e-9) % np.power(self.rho, -0.1) * np.power(i, -0.1)

        if self.w_flag == 1:
            self.w = np.sign(self.w) * np.where((np.abs(self.w) - self.t) > 0, np.abs(self.w) - self.t, 0)
            print('final true loss=', self.objective.get_arrogate_loss(torch.from_numpy(self.w).double()))
        return self.w

    def final_res(self):
        return super(smoothADMMmethod, self).final_res()
