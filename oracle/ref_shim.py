"""Loader for the *unmodified* reference (RufengXiao/ADMM-for-rank-based-loss) with the
two-line `get_opt` shim described in SURVEY.md Appendix B.

TEST INFRASTRUCTURE ONLY.  This module is used in the build container (where
`/root/reference` exists) to (a) validate the restated oracle in `oracle/rbl_oracle.py`
and (b) generate the committed golden vectors under `tests/golden/` (see
`oracle/gen_golden.py`).  Nothing on the product path, in `bench.py`, `smoke()` or the
`-m gpu` tests imports it: `/root/reference` does not exist on the GPU box.

Why a shim: the shipped reference unpacks two values from `solver.get_opt(...)`
(src/optim/algorithms.py:97,101) while `PAV_solver.get_opt` (src/util/pav.py:178) and
`PAV_solver_CPT.get_opt` (src/util/PAV_cpt.py:293) return one array, so every z-step
raises ValueError.  We wrap both to return `(res, None)` without touching reference files.
"""
import contextlib
import io
import os
import sys

REF_ROOT = os.environ.get("RBL_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "optim"))


def load(fista_dtype=None):
    """Import the reference's hot-path modules; returns a namespace object.

    fista_dtype: None -> as shipped (float32 FISTA, algorithms.py:201);
                 torch.float64 -> the "reference float64" l1 path (fast_lasso.py:22-26).
    """
    global _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    if _loaded is None:
        # The reference's top-level package is called `src`; make sure ours (the
        # drop-in mirror) is not shadowing it in this process.
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.path.insert(0, REF_ROOT)
        try:
            import src.util.pav as pav
            import src.util.PAV_cpt as pav_cpt

            _orig = pav.PAV_solver.get_opt
            _orig_cpt = pav_cpt.PAV_solver_CPT.get_opt

            def get_opt(self, maxiter=10000):
                return _orig(self, maxiter), None

            def get_opt_cpt(self):
                return _orig_cpt(self), None

            pav.PAV_solver.get_opt = get_opt
            pav_cpt.PAV_solver_CPT.get_opt = get_opt_cpt
            pav.PAV_solver._orig_get_opt = _orig
            pav_cpt.PAV_solver_CPT._orig_get_opt = _orig_cpt

            import src.optim.algorithms as algorithms
            import src.optim.objective as objective
            import src.util.individual_solver as individual_solver
            import src.util.fast_lasso as fast_lasso
            import src.util.w_LBFGS as w_lbfgs
            import src.util.load_data as load_data
        finally:
            sys.path.remove(REF_ROOT)

        class NS:
            pass

        ns = NS()
        ns.pav, ns.pav_cpt = pav, pav_cpt
        ns.algorithms, ns.objective = algorithms, objective
        ns.individual_solver, ns.fast_lasso = individual_solver, fast_lasso
        ns.w_lbfgs, ns.load_data = w_lbfgs, load_data
        ns._orig_fista = fast_lasso.FISTA
        _loaded = ns
    ns = _loaded

    orig = ns._orig_fista

    def fista(beta, X, y, lam, L, eta, tol=1e-4, max_iter=5000, dtype=torch.float32):
        dt = fista_dtype if fista_dtype is not None else dtype
        prev = torch.get_default_dtype()
        try:
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return orig(beta, X, y, lam, L, eta, tol=tol, max_iter=max_iter, dtype=dt)
        finally:
            torch.set_default_dtype(prev)  # undo fast_lasso.py:23-26 side effect

    ns.algorithms.FISTA = fista
    ns.FISTA = fista
    return ns


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
