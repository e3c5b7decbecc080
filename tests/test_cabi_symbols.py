"""CPU: the C-ABI library loads and exports every entry point include/dxb200.h declares (no compute calls)."""
import ctypes
import os
import re

from diffxpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dxb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert "dx_nb_fit_grouped" in names and "dx_group_suffstats" in names and "dx_two_group_tests" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libdxb200.so does not export %s" % n


def test_binding_covers_header():
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS)
    assert set(_declared()) == bound


def test_loader_and_version():
    lib = _lib.load()
    assert lib.dx_version() >= 100
    assert ctypes.sizeof(_lib.DxFitOpts) == 48


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    opts = _lib.DxFitOpts(max_steps=10, update_b_freq=5, train_loc=1, train_scale=1, init_a=1, init_b=1,
                          newton_max_iter=10, reserved=0, lltol=1e-10, newton_xtol=1e-8)
    rc = lib.dx_nb_fit_grouped(10, 0, 2, 1, None, None, None, None, None, None, 0, None, None, None, None, None,
                               ctypes.byref(opts), None, None, None, None, None, None, None, None)
    assert rc == 1 and b"n_groups" in lib.dx_last_error()
    rc = lib.dx_lrt(5, None, None, 1, None, None, None)
    assert rc == 1


def test_new_entry_points_validate_arguments_and_size_queries_run_on_the_host():
    lib = _lib.load()
    nbytes = ctypes.c_int64(-1)
    assert lib.dx_bh_work_bytes(20000, ctypes.byref(nbytes)) == 0
    assert nbytes.value >= 20 * 20000 + 8 * 20 and nbytes.value % 4 == 0        # 20 n + chunk minima + bucket tables
    small = ctypes.c_int64(-1)
    assert lib.dx_bh_work_bytes(0, ctypes.byref(small)) == 0 and 0 < small.value < nbytes.value
    assert lib.dx_bh_work_bytes(-1, ctypes.byref(small)) == 1
    # K_pack / packed fit / long-list rank test / peer gather: bad arguments come back as DX_ERR_INVALID_VALUE, never a crash
    assert lib.dx_pack_overflow(10, 0, 0, None, None, None, None, None, None) == 1 and b"n_groups" in lib.dx_last_error()
    assert lib.dx_pack_overflow(10, 4, 0, None, None, None, None, None, None) == 1 and b"null" in lib.dx_last_error()
    opts = _lib.DxFitOpts(max_steps=10, update_b_freq=5, train_loc=1, train_scale=1, init_a=1, init_b=1,
                          newton_max_iter=10, reserved=0, lltol=1e-10, newton_xtol=1e-8)
    rc = lib.dx_nb_fit_grouped_packed(10, 300, 2, 1, None, None, None, None, None, None, 0, None, None, None, None, None, None,
                                      ctypes.byref(opts), None, None, None, None, None, None, None, None)
    assert rc == 1 and b"n_groups" in lib.dx_last_error()
    rc = lib.dx_group_pair_tests_big(10, 3, 0, 0, 5.0, 5.0, None, None, None, None, None, None, None, None, None, None, 0, 0, None)
    assert rc == 1
    rc = lib.dx_peer_allgather_f64(10, None, None, 2, 0, 1024, None, None)
    assert rc == 1
    rc = lib.dx_exchange_open(None, None)
    assert rc == 1
