"""ctypes binding of the C ABI in include/bnmtf_b200.h.  Nothing here computes:
every numeric entry point goes to libbnmtf_b200.so, and there is no fallback --
a missing library or a missing CUDA device raises."""
import ctypes as C
import os

import numpy as np

from . import _build

GIBBS, ICM, VB, NP = 0, 1, 2, 3
SIDE_U, SIDE_V = 0, 1
F_VALUE, F_VAR, F_MU, F_TAU = 0, 1, 2, 3
NSTATS = 16
ST_RSS, ST_SUM_E, ST_SUM_RCE, ST_TAU, ST_BETA_S, ST_ESD, ST_ELBO, ST_IDIV, ST_LOGTAU, ST_ALPHA_S = range(10)

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_h = C.c_void_p

# name -> (restype, argtypes); mirrors include/bnmtf_b200.h declaration by declaration
SUM_FACTOR_U, SUM_FACTOR_V, SUM_S, SUM_LAMBDA_U, SUM_LAMBDA_V, SUM_TAU = range(6)
PRED_CURRENT, PRED_SUMMARY = 0, 1

SIGNATURES = {
    "bnmtf_last_error": (C.c_char_p, []),
    "bnmtf_device_count": (C.c_int, []),
    "bnmtf_abi_version": (C.c_int, []),
    "bnmtf_nmf_create": (C.c_int, [C.POINTER(_h), C.c_int, C.c_int64, C.c_int64, C.c_int, _dp, _u8p,
                                   C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "bnmtf_nmf_create_dev": (C.c_int, [C.POINTER(_h), C.c_int, C.c_int64, C.c_int64, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "bnmtf_nmf_layout_stats": (C.c_int, [_h, _i64p]),
    "bnmtf_nmf_pack": (C.c_int, [_h, _i64p]),
    "bnmtf_nmf_destroy": (C.c_int, [_h]),
    "bnmtf_nmf_data_stats": (C.c_int, [_h, _dp]),
    "bnmtf_nmf_set_data_stats": (C.c_int, [_h, C.c_double, C.c_double]),
    "bnmtf_nccl_unique_id": (C.c_int, [_u8p]),
    "bnmtf_nmf_set_comm": (C.c_int, [_h, C.c_int, C.c_int, _u8p]),
    "bnmtf_nmf_peer_export": (C.c_int, [_h, _u8p, C.c_int64]),
    "bnmtf_nmf_peer_import": (C.c_int, [_h, _u8p, C.c_int64]),
    "bnmtf_nmf_set_hyper": (C.c_int, [_h, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]),
    "bnmtf_nmf_set_factor": (C.c_int, [_h, C.c_int, C.c_int, _dp]),
    "bnmtf_nmf_get_factor": (C.c_int, [_h, C.c_int, C.c_int, _dp]),
    "bnmtf_nmf_set_scalars": (C.c_int, [_h, C.c_double, _dp]),
    "bnmtf_nmf_get_scalars": (C.c_int, [_h, _dp, _dp]),
    "bnmtf_nmf_set_vb_scalars": (C.c_int, [_h, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp]),
    "bnmtf_nmf_get_vb_scalars": (C.c_int, [_h, _dp, _dp]),
    "bnmtf_nmf_init_factors": (C.c_int, [_h, C.c_int, C.c_int, C.c_uint64]),
    "bnmtf_nmf_column_params": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "bnmtf_nmf_stats": (C.c_int, [_h, C.c_int, _dp]),
    "bnmtf_nmf_run": (C.c_int, [_h, C.c_int, C.c_int, C.c_uint64, _dp, _dp, _dp, _dp, _dp]),
    "bnmtf_nmf_launch_count": (C.c_int64, [_h]),
    "bnmtf_nmf_layout_info": (C.c_int, [_h, _i64p]),
    "bnmtf_nmf_sweep_times": (C.c_int, [_h, _dp]),
    "bnmtf_nmf_time_sweeps": (C.c_int, [_h, C.c_int]),
    "bnmtf_nmtf_create": (C.c_int, [C.POINTER(_h), C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, _dp, _u8p,
                                    C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "bnmtf_nmtf_create_dev": (C.c_int, [C.POINTER(_h), C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "bnmtf_nmtf_set_hyper": (C.c_int, [_h, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]),
    "bnmtf_nmtf_set_S": (C.c_int, [_h, C.c_int, _dp]),
    "bnmtf_nmtf_get_S": (C.c_int, [_h, C.c_int, _dp]),
    "bnmtf_nmtf_set_scalars": (C.c_int, [_h, C.c_double, _dp, _dp]),
    "bnmtf_nmtf_get_scalars": (C.c_int, [_h, _dp, _dp, _dp]),
    "bnmtf_nmtf_init": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64]),
    "bnmtf_nmtf_column_params": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "bnmtf_nmtf_stats": (C.c_int, [_h, C.c_int, _dp]),
    "bnmtf_nmtf_run": (C.c_int, [_h, C.c_int, C.c_int, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "bnmtf_nmtf_set_vb_scalars": (C.c_int, [_h, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]),
    "bnmtf_nmtf_get_vb_scalars": (C.c_int, [_h, _dp, _dp, _dp]),
    "bnmtf_nmtf_vb_update_params": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "bnmtf_nmtf_vb_stats": (C.c_int, [_h, _dp]),
    "bnmtf_nmtf_vb_run": (C.c_int, [_h, C.c_int, _dp, _dp]),
    "bnmtf_nmtf_np_update": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, _dp]),
    "bnmtf_nmtf_np_run": (C.c_int, [_h, C.c_int, _dp, _dp]),
    "bnmtf_kmeans_assign": (C.c_int, [_h, C.c_int, C.c_int, _dp, _u8p, C.POINTER(C.c_int32), _dp]),
    "bnmtf_kmeans_update": (C.c_int, [_h, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _dp, _u8p]),
    "bnmtf_kmeans_minmax": (C.c_int, [_h, C.c_int, _dp, _dp]),
    "bnmtf_nmf_set_summary": (C.c_int, [_h, C.c_int, C.c_int]),
    "bnmtf_nmf_get_summary": (C.c_int, [_h, C.c_int, _dp, C.POINTER(C.c_int64)]),
    "bnmtf_nmf_predict": (C.c_int, [_h, C.c_int, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _dp, _dp]),
    "bnmtf_batch_run": (C.c_int, [C.c_int, C.POINTER(_h), C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(_dp), _dp,
                                  C.POINTER(C.c_int64)]),
    "bnmtf_pool_reserve": (C.c_int, [C.c_int, C.c_int64]),
    "bnmtf_tn_draw": (C.c_int, [C.c_int, C.c_uint64, C.c_int64, _dp, _dp, _dp]),
    "bnmtf_tn_moments": (C.c_int, [C.c_int, C.c_int64, _dp, _dp, _dp, _dp]),
    "bnmtf_gamma_draw": (C.c_int, [C.c_int, C.c_uint64, C.c_int64, _dp, _dp, _dp]),
    "bnmtf_exp_draw": (C.c_int, [C.c_int, C.c_uint64, C.c_int64, _dp, _dp]),
}

_lib = None


def lib_path():
    # BNMTF_LIB: timing experiments only (an instrumented build of the same sources, e.g. build/libbnmtf_prof.so)
    return os.environ.get("BNMTF_LIB") or _build.LIB


def load():
    """Load libbnmtf_b200.so (building it if nvcc is available and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001
            raise RuntimeError("libbnmtf_b200.so is missing and could not be built (%s). "
                               "There is no CPU fallback: run `python -m bnmtf_ard_b200._build`." % exc)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("bnmtf_b200: " + load().bnmtf_last_error().decode("utf-8", "replace"))


def require_device():
    lib = load()
    if lib.bnmtf_device_count() < 1:
        raise RuntimeError("bnmtf_b200: no CUDA device visible. The models run on a B200 (sm_100a) only; "
                           "there is no CPU fallback.")
    return lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


def current_device():
    """Device for this process: LOCAL_RANK under torchrun, else BNMTF_DEVICE, else 0."""
    for key in ("BNMTF_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0
