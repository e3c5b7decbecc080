"""
ctypes binding of libdxb200.so (the C ABI declared in include/dxb200.h).

There is no CPU fallback: if the CUDA library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C diffxpy_b200/csrc``) every compute entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DXB200_LIB") or os.path.join(_HERE, "libdxb200.so")   # env override: tuning builds

DX_HIST_BINS = 64
DX_MAX_GROUPS = 255
DX_SKIP_GROUP = 255
DX_MAX_COEF = 32

DX_SHARD_INTEGER_COUNTS = 1
DX_GROUPS_FIT_U16 = 2
DX_SHARD_NO_OVERFLOW = 4
DX_OPT_SCALE_CONST = 1
DX_IPC_HANDLE_BYTES = 64
DX_INIT_GIVEN, DX_INIT_STANDARD, DX_INIT_CLOSED_FORM, DX_INIT_ALL_ZERO = 0, 1, 2, 3

FLAG_CONVERGED, FLAG_SINGULAR_FIM, FLAG_SCALE_FROZEN, FLAG_NONFINITE, FLAG_ALL_ZERO, FLAG_RANK_OVERFLOW = 1, 2, 4, 8, 16, 32

TG_MEAN0, TG_MEAN1, TG_VAR0, TG_VAR1, TG_PVAL_T, TG_PVAL_RANK, TG_NCOL = 0, 1, 2, 3, 4, 5, 6


class DxFitOpts(C.Structure):
    """struct dx_fit_opts (include/dxb200.h)."""
    _fields_ = [
        ("max_steps", C.c_int32), ("update_b_freq", C.c_int32), ("train_loc", C.c_int32), ("train_scale", C.c_int32),
        ("init_a", C.c_int32), ("init_b", C.c_int32), ("newton_max_iter", C.c_int32), ("reserved", C.c_int32),
        ("lltol", C.c_double), ("newton_xtol", C.c_double),
    ]


class DxError(RuntimeError):
    pass


_P = C.c_void_p
_I64, _I32, _F64 = C.c_int64, C.c_int32, C.c_double

# name -> argtypes; every function returns int
SIGNATURES = {
    "dx_device_sm_count": [C.POINTER(C.c_int)],
    "dx_fit_percell_counters": [C.POINTER(C.c_uint64), _I32],
    "dx_group_suffstats": [_I64, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P],
    "dx_group_suffstats_ex": [_I64, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _I32, _P],
    "dx_group_suffstats_pieces": [_I64, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _I64, _P, _I64, _I32, _P],
    "dx_upload": [_P, _P, _I64, _I32, _P],
    "dx_csr_to_csc_work_bytes": [_I64, _I32, C.POINTER(C.c_int64)],
    "dx_csr_to_csc": [_I64, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P],
    "dx_csr_to_csc_staged_work_bytes": [_I64, _I32, _I64, C.POINTER(C.c_int64)],
    "dx_csr_to_csc_staged": [_I64, _I64, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _I64, _P],
    "dx_gene_moments": [_I64, _P, _P, _P, _P, _P, _P, _P],
    "dx_nb_fit_grouped": [_I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, C.POINTER(DxFitOpts),
                          _P, _P, _P, _P, _P, _P, _P, _P],
    "dx_pack_overflow": [_I64, _I32, _I32, _P, _P, _P, _P, _P, _P],
    "dx_nb_fit_grouped_packed": [_I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P,
                                 C.POINTER(DxFitOpts), _P, _P, _P, _P, _P, _P, _P, _P],
    "dx_nb_fit_wald_grouped": [_I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P,
                               C.POINTER(DxFitOpts), _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P,
                               _P, _I32, _I32, _I64, _I64, _P],
    "dx_bh_qvalues_exchange": [_I64, _P, _I32, _I64, _P, _P, _P, _P],
    "dx_host_release": [],
    "dx_nb_fit_percell": [_I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, C.POINTER(DxFitOpts),
                          _P, _P, _P, _P, _P, _P, _P, _P],
    "dx_wald_test": [_I64, _P, _P, _P, _P, _P],
    "dx_wald_from_fit": [_I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P],
    "dx_wald_chisq": [_I64, _I32, _P, _P, _P, _P, _P, _P],
    "dx_lrt": [_I64, _P, _P, _I32, _P, _P, _P],
    "dx_two_coef_z": [_I64, _P, _P, _P, _P, _P, _P],
    "dx_t_test_moments": [_I64, _P, _P, _P, _P, _F64, _F64, _P, _P],
    "dx_bh_work_bytes": [_I64, C.POINTER(C.c_int64)],
    "dx_bh_qvalues": [_I64, _P, _P, _P, _P],
    "dx_bh_qvalues_sorted": [_I64, _P, _P, _P, _P, _P],
    "dx_two_group_tests": [_I64, _F64, _F64, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P],
    "dx_group_pair_tests": [_I64, _I32, _I32, _I32, _F64, _F64, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P],
    "dx_group_pair_tests_big": [_I64, _I32, _I32, _I32, _F64, _F64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P],
    "dx_exchange_create": [_I64, C.POINTER(C.c_void_p), _P],
    "dx_exchange_open": [_P, C.POINTER(C.c_void_p)],
    "dx_exchange_close": [_P],
    "dx_exchange_destroy": [_P],
    "dx_exchange_status": [_P, C.POINTER(C.c_int32)],
    "dx_peer_allgather_f64": [_I64, _P, _P, _I32, _I32, _I64, _P, _P],
    "dx_wald_grouped_host": [_I64, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, C.POINTER(DxFitOpts),
                             _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
}
OTHER_SYMBOLS = ["dx_version", "dx_last_error"]

_lib = None


def load():
    """Load the shared library once; raise DxError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DxError(
            "libdxb200.so is not built (%s). Build it with `make -C diffxpy_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.dx_version.restype = C.c_int
    lib.dx_version.argtypes = []
    lib.dx_last_error.restype = C.c_char_p
    lib.dx_last_error.argtypes = []
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point and raise DxError on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.dx_last_error()
        raise DxError("%s failed (code %d): %s" % (name, rc, msg.decode() if msg else "?"))
    return rc


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise DxError("diffxpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    load()
