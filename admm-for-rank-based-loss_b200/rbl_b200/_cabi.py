"""ctypes binding of librbl_b200.so (include/rbl_b200.h).  This is the only way the host layer reaches
the GPU: there is no CPU fallback — a missing library or a non-sm_100 device raises."""
import ctypes
import os

from . import build as _build

_c = ctypes
_dp = _c.c_void_p  # device pointers travel as integers (tensor.data_ptr())

class RunStats(_c.Structure):
    """rbl_run_stats (include/rbl_b200.h)"""
    _fields_ = [("rho", _c.c_double), ("primal", _c.c_double), ("dual", _c.c_double), ("iters", _c.c_int32),
                ("converged", _c.c_int32), ("rho_is_pyfloat", _c.c_int32), ("nnz_last", _c.c_int32),
                ("last_sweeps", _c.c_int32), ("pad", _c.c_int32), ("fista_iters", _c.c_int64),
                ("fista_sweeps", _c.c_int64), ("sparse_dual", _c.c_int64), ("dense_dual", _c.c_int64),
                ("gathered", _c.c_int64), ("rows_read", _c.c_int64)]


_SIGNATURES = {
    "rbl_version": (_c.c_int, []),
    "rbl_last_error": (_c.c_char_p, []),
    "rbl_launch_count": (_c.c_int64, []),
    "rbl_create": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_int, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int32,
                              _c.c_int64]),
    "rbl_destroy": (_c.c_int, [_c.c_void_p]),
    "rbl_set_storage": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "rbl_set_pass_grid": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "rbl_info": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_int64)]),
    "rbl_build_design": (_c.c_int, [_c.c_void_p, _dp, _c.c_int64, _dp, _dp, _c.c_void_p]),
    "rbl_build_design_rows": (_c.c_int, [_c.c_void_p, _dp, _c.c_int64, _dp, _dp, _c.c_int64, _c.c_void_p]),
    "rbl_set_spectrum": (_c.c_int, [_c.c_void_p, _dp, _c.c_void_p]),
    "rbl_matvec": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_void_p]),
    "rbl_margins": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_double, _dp, _c.c_void_p]),
    "rbl_sort_margins": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_void_p]),
    "rbl_pav_prox": (_c.c_int, [_c.c_void_p, _c.c_int, _dp, _c.c_double, _dp, _c.c_void_p]),
    "rbl_admm_run": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int32,
                                _c.c_double, _c.c_double, _c.c_int64, _c.c_int32, _c.c_int64, _c.c_double,
                                _c.c_int32, _c.POINTER(RunStats)]),
    "rbl_admm_run_l2": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _dp,
                                   _dp, _dp, _c.c_void_p, _dp, _c.c_double, _c.c_int32, _c.c_int32, _c.c_double,
                                   _c.c_int32, _c.c_int64, _c.c_double, _c.POINTER(RunStats)]),
    "rbl_bind_scalars": (_c.c_int, [_c.c_void_p, _dp]),
    "rbl_sort_margins_near": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _dp, _c.c_void_p]),
    "rbl_sort_stats": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_int32)]),
    "rbl_sort_debug": (_c.c_int, [_c.c_void_p, _dp]),
    "rbl_sort_config": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "rbl_pav_config": (_c.c_int, [_c.c_void_p, _c.c_int, _c.POINTER(_c.c_int32)]),
    "rbl_prox_elementwise": (_c.c_int, [_c.c_void_p, _c.c_int, _dp, _dp, _c.c_int64, _c.c_double, _dp, _c.c_void_p]),
    "rbl_ehrm_candidate_sums": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_double, _c.c_double, _dp, _c.c_void_p]),
    "rbl_scatter_z": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_int, _c.c_double, _dp, _c.c_double, _dp, _dp,
                                 _c.c_void_p]),
    "rbl_scatter_active": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_int, _c.c_double, _dp, _c.c_double, _dp, _dp,
                                      _c.c_void_p]),
    "rbl_grad_pass": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _dp, _c.c_int64, _dp, _c.c_void_p]),
    "rbl_gather_only": (_c.c_int, [_c.c_void_p, _dp, _c.c_void_p]),
    "rbl_active_count": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_int32), _c.c_void_p]),
    "rbl_fused_pass": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _dp, _dp, _c.c_void_p]),
    "rbl_fista_config": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_float)]),
    "rbl_fista_begin": (_c.c_int, [_c.c_void_p, _dp, _c.c_double, _c.c_int, _c.c_float, _c.c_double, _c.c_int,
                                   _c.c_void_p]),
    "rbl_fista_pass": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_void_p]),
    "rbl_fista_bind_red": (_c.c_int, [_c.c_void_p, _dp]),
    "rbl_fista_update": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "rbl_fista_steps": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_int, _c.c_void_p]),
    "rbl_fista_poll": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_double)]),
    "rbl_fista_result": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_void_p]),
    "rbl_gram_build": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_void_p]),
    "rbl_gram_accumulate": (_c.c_int, [_c.c_void_p, _dp, _c.c_int64, _c.c_int, _dp, _c.c_void_p]),
    "rbl_gram_fista_begin": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_double, _c.c_int, _c.c_float, _c.c_double,
                                        _c.c_int, _c.c_void_p]),
    "rbl_gram_fista_persistent_ok": (_c.c_int, [_c.c_void_p]),
    "rbl_gram_fista_run": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_double, _c.c_int, _c.c_float, _c.c_double,
                                      _c.c_int, _dp, _dp, _c.c_int, _c.c_void_p]),
    "rbl_gram_fista_steps": (_c.c_int, [_c.c_void_p, _dp, _c.c_int, _c.c_void_p]),
    "rbl_gram_fista_result": (_c.c_int, [_c.c_void_p, _dp, _c.c_void_p]),
    "rbl_lbfgs_gram": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_double, _c.c_double, _c.c_int, _c.c_double, _c.c_int,
                                  _c.c_void_p, _dp, _c.POINTER(_c.c_int32), _c.c_void_p]),
    "rbl_gram_eval_host": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "rbl_gram_eval": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _dp, _dp, _c.c_void_p]),
    "rbl_lasso_cd_gram": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _c.c_double, _c.c_double, _c.c_int, _dp, _dp,
                                     _c.c_void_p]),
    "rbl_h2d_pageable": (_c.c_int, [_c.c_int, _dp, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p]),
    "rbl_cpt_weights": (_c.c_int, [_c.c_int64, _c.c_int, _c.POINTER(_c.c_double)]),
    "rbl_dual_pass": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _c.c_double, _c.c_int, _c.c_int,
                                 _dp, _dp, _c.c_void_p]),
    "rbl_build_transpose": (_c.c_int, [_c.c_void_p, _dp, _dp, _c.c_void_p]),
    "rbl_batch_create": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "rbl_fista_batch_begin": (_c.c_int, [_c.c_void_p, _c.c_int, _dp, _c.POINTER(_c.c_double), _c.POINTER(_c.c_int32),
                                         _c.c_float, _c.c_double, _c.c_int, _c.c_void_p]),
    "rbl_fista_batch_steps": (_c.c_int, [_c.c_void_p, _c.c_int, _dp, _dp, _c.c_int, _c.c_void_p]),
    "rbl_fista_batch_poll": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.POINTER(_c.c_int32),
                                        _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _c.POINTER(_c.c_double)]),
    "rbl_fista_batch_result": (_c.c_int, [_c.c_void_p, _c.c_int, _dp, _dp, _c.c_void_p]),
    "rbl_dual_update": (_c.c_int, [_c.c_void_p, _dp, _dp, _dp, _dp, _c.c_int, _dp, _c.c_double, _dp, _dp, _dp,
                                   _c.c_void_p]),
    "rbl_objective": (_c.c_int, [_c.c_void_p, _c.c_int, _dp, _dp, _dp, _dp, _c.c_void_p]),
    "rbl_standardize_scratch_bytes": (_c.c_int, [_c.c_int, _c.c_int64, _c.POINTER(_c.c_int64)]),
    "rbl_standardize_columns": (_c.c_int, [_c.c_int, _dp, _c.c_int64, _c.c_int64, _c.c_int64, _dp, _dp, _dp,
                                           _c.c_void_p]),
    "rbl_gather_rows": (_c.c_int, [_c.c_int, _dp, _c.c_int64, _dp, _c.c_int64, _c.c_int64, _dp, _c.c_int64,
                                   _c.c_void_p]),
    "rbl_metrics_scratch_bytes": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_int64)]),
    "rbl_test_metrics": (_c.c_int, [_c.c_int, _dp, _c.c_int64, _c.c_int64, _c.c_int64, _dp, _dp, _dp, _c.c_int,
                                    _c.c_double, _dp, _dp, _c.c_void_p]),
}

_lib = None
ABI_VERSION = 2  # RBL_ABI_VERSION in include/rbl_b200.h


class RblError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load():
    """Load the shared library; it is (re)built first when it is absent, or older than its sources and nvcc is at
    hand (a stale binary after editing csrc/ would otherwise be used silently)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RblError(f"librbl_b200.so is missing at {path} and could not be built ({e}); "
                           "run `python __graft_entry__.py build`. There is no CPU fallback.") from e
    elif _build.stale() and _build.have_nvcc():
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RblError(f"librbl_b200.so is older than csrc/ and the rebuild failed ({e})") from e
    lib = ctypes.CDLL(path)
    try:
        got = lib.rbl_version()
    except AttributeError as e:
        raise RblError(f"{path} does not export rbl_version: not a librbl_b200 build") from e
    if got != ABI_VERSION:
        raise RblError(f"librbl_b200.so ABI version {got} != {ABI_VERSION} expected by this package; rebuild with "
                       "`python __graft_entry__.py build`")
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RblError(f"librbl_b200.so does not export {name} (stale build?); rebuild with "
                           "`python __graft_entry__.py build`") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc):
    if rc != 0:
        msg = load().rbl_last_error()
        raise RblError(f"librbl_b200 error {rc}: {msg.decode() if msg else ''}")
