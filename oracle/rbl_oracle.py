"""CPU oracle: a numpy/C restatement of the reference's ADMM hot path.

TEST INFRASTRUCTURE ONLY — only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module.  The product path
(`admm-for-rank-based-loss_b200/`) never imports anything under `oracle/`; it fails loudly
if the CUDA library is missing.

Every function cites the reference file:line (paths under the reference tree) it follows.
The reference is pure Python (no native code), so there is no `oracle/_ref` binary; the
restatement is pinned against outputs of the reference itself, generated in the build
container by `oracle/gen_golden.py` (shimmed import) and committed
under `tests/golden/`.  The reference has no tests or golden vectors of its own: parity is
"unpinned" by the reference; the only numeric artefact it ships is the xlsx table whose
iteration-0 objective 0.6931471805674658 reproduces bit-exactly (tests/test_oracle.py).

Deliberate deviations from the shipped reference:
  * z-step uses an exact stack PAV + bracketed Newton (machine precision) instead of the
    sweep PAV with a global-stop damped Newton (pav.py:93-178, individual_solver.py:90-109);
    the isotonic prox is unique, measured agreement <= ~1e-10 abs for BCE.
  * hinge block prox is the closed form, not the early-exit bisection
    (individual_solver.py:15-42), which is inexact at small rho (reference defect).
  * FISTA runs in float64 (`dtype=torch.float64`, fast_lasso.py:22-26) unless fp32 is asked
    for; scalar promotion follows numpy >= 2 (NEP 50), which is what this image runs:
    `L_cur` and `lam/L_cur` are float32 scalars (algorithms.py:199-200 pass np.float32).
  * sort is `kind="stable"` (reference: unstable np.argsort, algorithms.py:92; identical on
    tie-free keys).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LOSS_IDS = {"binary_cross_entropy": 0, "hinge": 1}


def build_c(force=False):
    """gcc-compile oracle/pav_oracle.c -> oracle/_build/libpav_oracle.so (git-ignored)."""
    src = os.path.join(_HERE, "pav_oracle.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "libpav_oracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        dp = ctypes.POINTER(ctypes.c_double)
        lib.rbl_oracle_prox.restype = ctypes.c_double
        lib.rbl_oracle_prox.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        lib.rbl_oracle_pav.restype = ctypes.c_int64
        lib.rbl_oracle_pav.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_double, dp]
        lib.rbl_oracle_prox_vec.restype = None
        lib.rbl_oracle_prox_vec.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, ctypes.c_double, dp]
        lib.rbl_oracle_lasso_cd.restype = ctypes.c_int
        lib.rbl_oracle_lasso_cd.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_int, dp, ctypes.POINTER(ctypes.c_double)]
        _LIB = lib
    return _LIB


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ---------------------------------------------------------------------------------------
# spectra — src/optim/objective.py:97-187
# ---------------------------------------------------------------------------------------
def _distort(p, gamma):  # objective.py:148-150
    return p ** gamma / ((p ** gamma + (1 - p) ** gamma) ** (1 / gamma))


def spectrum(name, n, args=None):
    """sigma in ascending-rank order (objective.py:166-187).  For 'ehrm' returns (a, b)."""
    i = np.arange(n, dtype=np.float64)
    if name == "erm":  # :97-98
        return np.ones(n) / n
    if name == "ehrm":  # :153-164 (python loop over scalars in the reference)
        a = np.array([_distort((k + 1) / n, 0.69) - _distort(k / n, 0.69) for k in range(n)])
        b = np.array([_distort((n - k) / n, 0.61) - _distort((n - k - 1) / n, 0.61) for k in range(n)])
        return a, b
    if args is None:
        raise ValueError("args for framework is None!")
    if name == "extremile":  # :101-105
        r = args[0]
        return ((i + 1) ** r - i ** r) / (n ** r)
    if name == "superquantile":  # :108-117
        q = args[0]
        w = np.zeros(n)
        idx = math.floor(n * q)
        frac = 1 - (n - idx - 1) / (n * (1 - q))
        if frac > 1e-12:
            w[idx] = frac
            w[idx + 1:] = 1 / (n * (1 - q))
        else:
            w[idx:] = 1 / (n - idx)
        return w
    if name == "esrm":  # :120-123
        rho = args[0]
        upper = np.exp(rho * ((i + 1) / n))
        lower = np.exp(rho * (i / n))
        return math.exp(-rho) * (upper - lower) / (1 - math.exp(-rho))
    if name == "aorr":  # :126-136
        qlow, qup = args[0], args[1]
        w = np.zeros(n)
        lo = math.floor(n * qlow)
        up = math.floor(n * qup)
        frac = 1 - (up - lo - 1) / (n * (qup - qlow))
        if frac > 1e-12:
            w[lo] = frac
            w[lo + 1:up] = 1 / (n * (qup - qlow))
        else:
            w[lo:up] = 1 / (up - lo)
        return w
    if name == "aorr_dc":  # :139-145
        k, m = args[0], args[1]
        if k <= m:
            raise ValueError("need args[0] > args[1]!")
        w = np.zeros(n)
        w[m + 1:k] = 1 / (k - m)
        w[k + 1] = 1 - (k - m - 1) / (k - m)
        return w
    raise ValueError(f"Unrecognized framework '{name}'!")


# ---------------------------------------------------------------------------------------
# losses / objective — objective.py:11-24, 71-87
# ---------------------------------------------------------------------------------------
def log1pexp(u):  # individual_solver.py:52-57
    u = np.asarray(u, dtype=np.float64)
    return np.where(u > 0, u + np.log1p(np.exp(-np.abs(u))), np.log1p(np.exp(-np.abs(u))))


def margin_loss(loss, u):
    """individual loss as a function of the margin u = D w = -y * (X w)."""
    if loss == "binary_cross_entropy":
        return log1pexp(u)
    if loss == "hinge":
        return np.maximum(1.0 + u, 0.0)
    raise ValueError(f"Unrecognized loss '{loss}'!")


def objective(D, w, sigma, loss, l2_reg=None, l1_reg=None, lossB=None, include_reg=True):
    """get_arrogate_loss (objective.py:71-87) in terms of D = -y*X.

    Note the reference's EHRM branch uses `alphas` on both sides of lossB (:76), so the value
    is sum(alphas * sorted losses) in every case; `lossB` is accepted and ignored on purpose.
    """
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    losses = np.sort(margin_loss(loss, D @ w))
    risk = float(np.dot(sigma, losses))
    if l2_reg and include_reg:
        risk += 0.5 * l2_reg * float(np.sum(w ** 2))
    if l1_reg and include_reg:
        risk += 0.5 * l1_reg * float(np.sum(np.abs(w)))
    return risk


# ---------------------------------------------------------------------------------------
# z-step — algorithms.py:88-106, pav.py:54-178, individual_solver.py:90-130, PAV_cpt.py
# ---------------------------------------------------------------------------------------
def prox_vec(loss, sigma, m, rho):
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    m = np.ascontiguousarray(m, dtype=np.float64)
    out = np.empty_like(m)
    _lib().rbl_oracle_prox_vec(LOSS_IDS[loss], m.size, _dptr(sigma), _dptr(m), float(rho), _dptr(out))
    return out


def pav_prox(loss, sigma, m_sorted, rho, clip=None, return_blocks=False):
    """exact isotonic prox of sorted margins (pav.py:93-178 fixed point); clip = EHRM's B."""
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    m_sorted = np.ascontiguousarray(m_sorted, dtype=np.float64)
    out = np.empty_like(m_sorted)
    nb = _lib().rbl_oracle_pav(LOSS_IDS[loss], m_sorted.size, _dptr(sigma), _dptr(m_sorted), float(rho),
                               0 if clip is None else 1, 0.0 if clip is None else float(clip), _dptr(out))
    return (out, nb) if return_blocks else out


def pav_prox_minmax(loss, sigma, m_sorted, rho):
    """Brute-force min-max formula z_i = max_{l<=i} min_{r>=i} v(l, r) (small n only)."""
    n = len(m_sorted)
    lib = _lib()
    cs = np.concatenate([[0.0], np.cumsum(sigma)])
    cm = np.concatenate([[0.0], np.cumsum(m_sorted)])
    v = np.full((n, n), np.nan)
    for l in range(n):
        for r in range(l, n):
            c = r - l + 1
            v[l, r] = lib.rbl_oracle_prox(LOSS_IDS[loss], (cs[r + 1] - cs[l]) / c, (cm[r + 1] - cm[l]) / c, rho)
    z = np.empty(n)
    for i in range(n):
        z[i] = max(min(v[l, r] for r in range(i, n)) for l in range(i + 1))
    return z


def ehrm_candidate_sums(sigma_a, sigma_b, B, m_sorted, rho):
    """The two scalars the reference compares in PAV_solver_CPT.__init__ (PAV_cpt.py:203-226):
    fval1 = func_value(sigma_a, min(prox_a, B)), fval2 = func_value(sigma_b, max(prox_b, B))."""
    x1 = np.minimum(prox_vec("binary_cross_entropy", sigma_a, m_sorted, rho), B)
    x2 = np.maximum(prox_vec("binary_cross_entropy", sigma_b, m_sorted, rho), B)
    f1 = float(np.sum(sigma_a * log1pexp(x1)) + rho / 2 * np.dot(x1 - m_sorted, x1 - m_sorted))
    f2 = float(np.sum(sigma_b * log1pexp(x2)) + rho / 2 * np.dot(x2 - m_sorted, x2 - m_sorted))
    return f1, f2


def ehrm_pav(sigma_a, sigma_b, B, m_sorted, rho, return_choice=False):
    """PAV_solver_CPT(sigma_a, sigma_b, B, m_sorted, rho).get_opt() (PAV_cpt.py:169-293).  The reference picks
    between its two clipped candidates by comparing two SCALARS (:222-226 — `fval1 <= fval2` indexes the whole
    vector), so the start is candidate 1 = min(prox_a, B) everywhere or candidate 2 = max(prox_b, B) everywhere,
    and the sweep that follows (:229-288) pools the winner: the result is min(B, isotonic prox with sigma_a) or
    max(B, isotonic prox with sigma_b) (pinned against the reference's outputs in tests/golden/zstep_ehrm.npz)."""
    f1, f2 = ehrm_candidate_sums(sigma_a, sigma_b, B, m_sorted, rho)
    if f1 <= f2:
        zs = np.minimum(pav_prox("binary_cross_entropy", sigma_a, m_sorted, rho), B)
    else:
        zs = pav_prox("binary_cross_entropy", sigma_b, m_sorted, rho, clip=B)
    return (zs, 1 if f1 <= f2 else 2) if return_choice else zs


def ehrm_sweep_literal(sigma_a, sigma_b, B, m_sorted, rho):
    """Small-n literal restatement of PAV_solver_CPT (PAV_cpt.py:203-293): per pass, block values of BOTH
    candidates from the unweighted block means (:122-123), the scalar choice between them re-made on the pooled
    blocks (:262-288), adjacent violators merged in chains (:236-246).  Exact block solves instead of the
    reference's tol-1e-4 Newton.  Returns (z, list of per-pass choices) — used to check that the choice made at
    element level is the one every later pass makes too."""
    blocks = [[i] for i in range(len(m_sorted))]
    choices = []
    while True:
        s1 = np.array([np.mean(sigma_a[b]) for b in blocks])
        s2 = np.array([np.mean(sigma_b[b]) for b in blocks])
        mm = np.array([np.mean(m_sorted[b]) for b in blocks])
        x1 = np.minimum(prox_vec("binary_cross_entropy", s1, mm, rho), B)
        x2 = prox_vec("binary_cross_entropy", s2, mm, rho)
        x2[x2 <= B] = B
        f1 = float(np.sum(s1 * log1pexp(x1)) + rho / 2 * np.dot(x1 - mm, x1 - mm))
        f2 = float(np.sum(s2 * log1pexp(x2)) + rho / 2 * np.dot(x2 - mm, x2 - mm))
        x = x1 if f1 <= f2 else x2
        choices.append(1 if f1 <= f2 else 2)
        new, i, viol = [], 0, False
        while i < len(blocks) - 1:
            if x[i] <= x[i + 1]:
                new.append(blocks[i])
            else:
                viol = True
                cur = list(blocks[i])
                while i < len(blocks) - 1 and x[i] > x[i + 1]:
                    cur = cur + blocks[i + 1]
                    i += 1
                new.append(cur)
            i += 1
        if len(blocks) >= 2 and x[-1] >= x[-2]:
            new.append(blocks[-1])
        if not viol:
            break
        blocks = new
    z = np.empty(len(m_sorted))
    for b, v in zip(blocks, x):
        z[b] = v
    return z, choices


def z_step(D, w, lam, rho, sigma, loss, B=None, sigma_b=None, return_all=False):
    """Optimizer.z_subproblem (algorithms.py:88-106)."""
    m = (D @ w.reshape(-1) - lam.reshape(-1) / rho)
    perm = np.argsort(m, kind="stable")
    ms = m[perm]
    if B is not None:  # EHRM, PAV_cpt.py:169-293
        zs = ehrm_pav(sigma, sigma_b, B, ms, rho)
    else:
        zs = pav_prox(loss, sigma, ms, rho)
    z = np.zeros(m.shape[0])
    z[perm] = zs
    if return_all:
        return z, m, perm, zs
    return z


# ---------------------------------------------------------------------------------------
# w-step, l1 — fast_lasso.py:15-69 (FISTA), algorithms.py:190-202
# ---------------------------------------------------------------------------------------
def soft_thr(x, alpha):  # fast_lasso.py:15-19
    return np.maximum(np.abs(x) - alpha, 0.0) * np.sign(x)


def fista(beta, X, y, lam, L=np.float32(17), eta=np.float32(2.5), tol=7e-5, max_iter=5000,
          dtype=np.float64, return_info=False):
    """FISTA for 0.5*||y - X b||^2 + lam*||b||_1 with backtracking (fast_lasso.py:22-69).

    Scalar types follow the reference call (algorithms.py:199-201) under numpy >= 2:
    L, eta are np.float32 so `L_cur = L_prev*(eta**i_k)` is a float32 product and
    `lam/L_cur` is a float32 quotient; tensor arithmetic is in `dtype`.
    """
    dt = np.dtype(dtype)
    X = np.asarray(X, dtype=dt)
    y = np.asarray(y, dtype=dt).reshape(-1)
    b = np.asarray(beta, dtype=dt).reshape(-1).copy()
    b_p = b.copy()
    b_prev = b.copy()
    t = dt.type(1.0)
    L_prev = L
    n_pass = 0
    k = 0
    for k in range(max_iter):
        r = y - X @ b_p
        drbp = r @ r
        g = X.T @ r
        n_pass += 2
        i_k = -1
        while True:
            i_k += 1
            L_cur = L_prev * (eta ** i_k)
            thr = lam / L_cur                      # float32 when L is float32 (NEP 50)
            b = soft_thr(b_p + g / dt.type(L_cur), dt.type(thr)).astype(dt, copy=False)
            diff = b - b_p
            rhs = dt.type(L_cur) * (diff @ diff) - dt.type(2.0) * (diff @ g)
            rr = y - X @ b
            n_pass += 1
            lhs = rr @ rr - drbp
            if not (lhs > rhs):
                break
        L_prev = L_cur
        tnext = (dt.type(1.0) + np.sqrt(dt.type(1.0) + dt.type(4.0) * t * t)) / dt.type(2.0)
        diff = b - b_prev
        t1 = (t - dt.type(1.0)) / tnext
        b_p = b + t1 * diff
        crit = np.sqrt(diff @ diff)
        if crit < tol:
            break
        t = tnext
        b_prev = b
    if return_info:
        return b, {"iters": k + 1, "passes": n_pass, "L": float(L_prev)}
    return b


def lasso_cd(X, y, alpha, tol=1e-8, max_iter=50000, return_info=False):
    """sklearn.linear_model.Lasso(alpha, tol, fit_intercept=False, max_iter).fit(X, y).coef_ as called at
    algorithms.py:195-196 — scikit-learn is third-party and un-vendored (README.md pins 1.2.2): restated from that
    release's published `enet_coordinate_descent` (cyclic coordinate descent, duality-gap stop; pav_oracle.c).
    Minimises 1/(2n) ||y - X w||^2 + alpha ||w||_1; pinned against the installed scikit-learn in tests/test_oracle."""
    X = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    n, d = X.shape
    w = np.zeros(d)
    gap = ctypes.c_double(0.0)
    sweeps = _lib().rbl_oracle_lasso_cd(n, d, _dptr(X), _dptr(y), float(alpha) * n, float(tol), int(max_iter),
                                        _dptr(w), ctypes.byref(gap))
    return (w, {"sweeps": sweeps, "gap": gap.value}) if return_info else w


# ---------------------------------------------------------------------------------------
# w-step, l2 — w_LBFGS.py:31-53
# ---------------------------------------------------------------------------------------
def w_step_l2(w0, z, lam, rho, D, reg, DTD=None, return_info=False):
    from scipy.optimize import minimize

    b = (z.reshape(-1) + lam.reshape(-1) / rho)
    DTb = D.T @ b

    def f(w):  # wl2_fun :31-37
        t = D @ w - b
        return 0.5 * rho * float(t @ t) + 0.5 * reg * float(w @ w)

    def g(w):  # wl2_fun_gradient :40-45
        if DTD is not None:
            return rho * (DTD @ w - DTb) + reg * w
        return rho * (D.T @ (D @ w - b)) + reg * w

    res = minimize(f, np.asarray(w0, dtype=np.float64).reshape(-1), jac=g, method="L-BFGS-B",
                   options={"maxiter": 1000})
    if return_info:
        return res.x, {"nit": res.nit, "nfev": res.nfev}
    return res.x


def w_step_smooth_l1(w0, z, lam, rho, D, reg, t, DTD=None, return_info=False):
    """smoothed-l1 w-step of smoothADMMmethod: w_LBFGS.py:11-28,54-62 (Huber-type smoothing of reg/2*|w|)."""
    from scipy.optimize import minimize

    b = (z.reshape(-1) + lam.reshape(-1) / rho)
    DTb = D.T @ b

    def f(w):  # wl1_fun_smooth :11-19
        tmp = D @ w - b
        res1 = 0.5 * rho * float(np.sum(np.square(tmp)))
        small = np.abs(w) <= t
        res2 = 0.5 * 0.5 * reg * float(np.sum(np.square(w[small]))) / t
        res2 += 0.5 * reg * float(np.sum(np.abs(w[~small]) - 0.5 * t))
        return res1 + res2

    def g(w):  # wl1_fun_smooth_gradient :22-28
        g1 = rho * ((DTD @ w - DTb) if DTD is not None else D.T @ (D @ w - b))
        small = np.abs(w) <= t
        return g1 + np.where(small, 0.5 * reg * w / t, 0.5 * reg * np.sign(w))

    res = minimize(f, np.asarray(w0, dtype=np.float64).reshape(-1), jac=g, method="L-BFGS-B",
                   options={"maxiter": 1000})
    if return_info:
        return res.x, {"nit": res.nit, "nfev": res.nfev}
    return res.x


# ---------------------------------------------------------------------------------------
# full loop — algorithms.py:20-75 (state), :119-164 (iteration), :190-216 (ADMMmethod)
# ---------------------------------------------------------------------------------------
class OracleADMM:
    def __init__(self, X, y, weight_function="erm", loss="binary_cross_entropy", l2_reg=None, l1_reg=None,
                 B=None, args=None, w0=None, max_iter=200, tol=1e-4, fista_dtype=np.float64, use_gram=True,
                 small_lasso=True):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y).reshape(-1, 1)
        self.D = np.ascontiguousarray(-y * X)  # :23
        self.n, self.d = X.shape
        self.DTD = self.D.T @ self.D if (use_gram and l1_reg is None) else None  # :24 (only the l2 path reads it)
        self.reg = l1_reg or l2_reg  # :30
        self.l1_reg, self.l2_reg = l1_reg, l2_reg
        self.lam = 0.1 * self.reg / self.n * np.ones(self.n)  # :32
        self.z = 0.1 * self.reg / self.n * np.ones(self.n)  # :34
        self.loss = loss
        self.w = (np.asarray(w0, dtype=np.float64).reshape(-1).copy() if w0 is not None
                  else 0.001 * self.reg / self.d / self.n * np.ones(self.d))  # :39-42
        self.tol, self.max_iter = tol, max_iter
        self.rho = 1e-4 if weight_function == "ehrm" else (2e-7 if weight_function in ("aorr", "aorr_dc") else 1e-5)  # :47-52
        self.w_flag = 1 if l1_reg is not None else 2  # :55-62
        if B is not None and weight_function != "ehrm":
            raise ValueError(f"Unrecognized weight_function '{weight_function}'! Options: ['ehrm']")
        self.B = B
        sig = spectrum(weight_function, self.n, args)
        self.sigma_a, self.sigma_b = sig if isinstance(sig, tuple) else (sig, sig)
        self.weight_function = weight_function
        self.fista_dtype = fista_dtype
        self.small_lasso = small_lasso
        self.primal = self.dual = float("inf")
        self.iters = 0
        self.passes = 0

    def objective(self, w=None):
        return objective(self.D, self.w if w is None else w, self.sigma_a, self.loss, self.l2_reg, self.l1_reg)

    def z_step(self):
        return z_step(self.D, self.w, self.lam, self.rho, self.sigma_a, self.loss,
                      B=self.B if self.weight_function == "ehrm" else None, sigma_b=self.sigma_b)

    def w_step(self):
        if self.w_flag == 1:  # algorithms.py:190-202
            b = self.z + self.lam / self.rho
            alpha = self.reg / (2 * self.rho * self.n)
            if self.n <= 500 and self.d <= 60 and self.small_lasso:  # :194-197 tiny-problem branch
                if self.small_lasso == "sklearn":  # the installed scikit-learn itself (cross-check only)
                    from sklearn.linear_model import Lasso

                    mdl = Lasso(alpha=alpha, tol=1e-8, fit_intercept=False, max_iter=50000, warm_start=True)
                    mdl.fit(X=self.D, y=b)
                    return mdl.coef_.reshape(-1).astype(np.float64)
                w, info = lasso_cd(self.D, b, alpha, tol=1e-8, max_iter=50000, return_info=True)
                self.last_cd_sweeps = info["sweeps"]
                return w
            w, info = fista(self.w, self.D, b, alpha * self.n, np.float32(17), np.float32(2.5), tol=7e-5,
                            max_iter=5000, dtype=self.fista_dtype, return_info=True)
            self.passes += info["passes"]
            self.last_fista_iters = info["iters"]
            self.last_fista_info = (info["iters"], info["passes"] - 2 * info["iters"], info["L"])  # (iters, trials, L)
            return w.astype(np.float64) if self.fista_dtype == np.float64 else w
        w, info = w_step_l2(self.w, self.z, self.lam, self.rho, self.D, self.reg, self.DTD, return_info=True)
        self.passes += 2 * info["nfev"]
        self.last_lbfgs_info = (int(info["nit"]), int(info["nfev"]))
        return w

    def step(self):
        """one Optimizer.main_loop body (:119-164); returns True on convergence."""
        self.z = self.z_step()
        pre_w = self.w.copy()
        self.w = self.w_step()
        Dw = self.D @ self.w
        self.lam = self.lam + self.rho * (self.z - Dw)  # :132
        self.primal = float(np.linalg.norm(self.z - Dw))  # :135
        self.dual = float(np.linalg.norm(self.w - pre_w))  # :136
        self.iters += 1
        if self.primal < self.tol and self.dual < self.tol:
            return True
        mult = 1.02 if self.primal > 1e-2 else 1.07  # :153-157 (the :148-152 branches are dead)
        # np.min returns np.float64: from the 2nd iteration on `lam/L_cur` in FISTA is a float64
        # quotient, in iteration 0 (python-float rho) it is a float32 one (NEP 50) — kept as is.
        self.rho = np.min((self.rho * mult, 217 * self.d))
        return False

    def main_loop(self):
        for _ in range(self.max_iter):
            if self.step():
                break
        return self.w


class OracleSmoothADMM(OracleADMM):
    """smoothADMMmethod (algorithms.py:223-263): Huber-smoothed l1 w-step, t schedule, final soft-threshold."""

    def __init__(self, *a, t=1, **kw):
        kw.setdefault("use_gram", True)
        super().__init__(*a, **kw)
        if self.DTD is None:
            self.DTD = self.D.T @ self.D
        self.t = t

    def w_step(self):
        if self.w_flag == 1:
            w, info = w_step_smooth_l1(self.w, self.z, self.lam, self.rho, self.D, self.reg, self.t, self.DTD,
                                       return_info=True)
            self.passes += 2 * info["nfev"]
            return w
        return super().w_step()

    def main_loop(self):
        for i in range(self.max_iter):
            if self.step():
                break
            if i >= 17:  # :254-255
                self.t = max(self.t * 0.9, 1e-9) % np.power(self.rho, -0.1) * np.power(i, -0.1)
        if self.w_flag == 1:  # :257-258
            self.w = np.sign(self.w) * np.where((np.abs(self.w) - self.t) > 0, np.abs(self.w) - self.t, 0)
        return self.w


# ---------------------------------------------------------------------------------------
# test-set metrics — the step after the path: calculate_acc.py:3-19, fair_metric.py:3-40
# ---------------------------------------------------------------------------------------
def class_probs(w, X):
    """sigmoid(X w), evaluated on the stable side of 0 (calculate_acc.py:5-8, fair_metric.py:4-7)."""
    p = np.asarray(X, dtype=np.float64) @ np.asarray(w, dtype=np.float64).reshape(-1)
    e = np.exp(-np.abs(p))
    return np.where(p >= 0, 1.0 / (1.0 + e), e / (e + 1.0))


def calculate_accuracy(w, X_test, y_test, threshold=0.5, loss="binary_cross_entropy"):
    """calculate_acc.py:3-19.  BCE: predict +1 iff sigmoid(x.w) >= threshold, else -1.  hinge: the reference maps
    (x.w >= 0) to {1, 0} and then rewrites the zeros to 1 (:14-15), so every prediction is +1 and the value is the
    share of positive labels — reproduced as shipped."""
    y = np.asarray(y_test).reshape(-1)
    if loss == "binary_cross_entropy":
        pred = np.where(class_probs(w, X_test) >= threshold, 1, -1)
    elif loss == "hinge":
        pred = np.ones(len(y), dtype=int)
    else:
        raise ValueError(f"loss '{loss}' is not supported! Options: ['binary_cross_entropy','hinge']")
    return float(np.mean(pred == y))


def confusion_by_group(w, X_test, label_test, group_test, threshold=0.5):
    """per group g in {0, 1}: [size, predicted positive, TP, FN, TN, FP] (fair_metric.py:11-24), plus
    sum(b), sum(b log b) for b = prob - y01 + 1 (:35-38)."""
    prob = class_probs(w, X_test)
    pred = prob >= threshold
    y01 = np.asarray(label_test).reshape(-1) == 1          # :9-10 (labels -1 -> 0)
    grp = np.asarray(group_test).reshape(-1)
    out = np.zeros((2, 6))
    for g in (0, 1):
        sel = grp == g
        out[g] = [sel.sum(), (pred & sel).sum(), (pred & y01 & sel).sum(), (~pred & y01 & sel).sum(),
                  (~pred & ~y01 & sel).sum(), (pred & ~y01 & sel).sum()]
    b = prob - y01 + 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        return out, float(b.sum()), float(np.sum(b * np.log(b)))


def statistics_from_counts(counts, sum_b, sum_blogb, n):
    """SPD, DI, EOD, AOD, TI, FNRD from the confusion counts (fair_metric.py:13-40).  Group 0 is the reference
    group ("G1"), group 1 the other ("G2").  TI = mean((b/mu) log(b/mu)) = sum(b log b)/(n mu) - log(mu).
    Empty classes divide by zero exactly as numpy does in the reference (nan / inf, no exception)."""
    c = np.asarray(counts, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        size, pp, tp, fn, tn, fp = (c[:, k] for k in range(6))
        gp = pp / size
        spd = gp[1] - gp[0]
        di = np.inf if gp[0] == 0 else gp[1] / gp[0]
        tpr, fpr, fnr = tp / (tp + fn), fp / (fp + tn), fn / (tp + fn)
        eod = tpr[1] - tpr[0]
        aod = 0.5 * (fpr[1] - fpr[0] + eod)
        mu = np.float64(sum_b) / n
        ti = np.float64(sum_blogb) / (n * mu) - np.log(mu)
        fnrd = fnr[1] - fnr[0]
    return float(spd), float(di), float(eod), float(aod), float(ti), float(fnrd)


def calculate_statistics(w, X_test, label_test, group_test, threshold=0.5):
    """fair_metric.py:3-40 -> (SPD, DI, EOD, AOD, TI, FNRD)."""
    counts, sb, sbl = confusion_by_group(w, X_test, label_test, group_test, threshold)
    return statistics_from_counts(counts, sb, sbl, len(np.asarray(label_test).reshape(-1)))
