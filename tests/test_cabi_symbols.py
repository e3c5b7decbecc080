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
