"""
Multiple-testing correction (/root/reference/diffxpy/testing/correction.py:5-24).  The reference calls
statsmodels.stats.multitest.multipletests(method="fdr_bh") on the non-NaN p-values; the same arithmetic
(p sorted ascending, q = p / (i/n), reverse running minimum, clip at 1, unsort) runs on the GPU here.
"""
import numpy as np
import torch

from .. import _lib
from ..device import cuda_device, ptr, stream_ptr

import ctypes as _C

BH_KERNEL_MAX = 1 << 24   # above this: device sort + dx_bh_qvalues_sorted


def bh_device(p):
    """Benjamini-Hochberg on a float64 CUDA vector; NaNs are excluded from n and stay NaN."""
    if p.shape[0] <= BH_KERNEL_MAX:
        p = p.contiguous()
        q = torch.empty_like(p)
        nbytes = _C.c_int64(0)
        _lib.call("dx_bh_work_bytes", p.shape[0], _C.byref(nbytes))
        work = torch.empty((nbytes.value + 7) // 8, dtype=torch.float64, device=p.device)
        _lib.call("dx_bh_qvalues", p.shape[0], ptr(p), ptr(q), ptr(work), stream_ptr())
        return q
    # large n: device sort (NaNs last) + one kernel for ranks, running minimum and the scatter back
    ps, order = torch.sort(p.contiguous())
    q = torch.empty_like(ps)
    work = torch.empty(ps.shape[0] // 1024 + 2, dtype=torch.float64, device=p.device)
    _lib.call("dx_bh_qvalues_sorted", ps.shape[0], ptr(ps), ptr(order), ptr(q), ptr(work), stream_ptr())
    return q


def correct(pvals, method="fdr_bh", alpha=0.05):
    """
    Performs multiple testing correction; only non-nan p-values are corrected (correction.py:15-22).

    :param pvals: uncorrected p-values. Must be 1-dimensional.
    :param method: "fdr_bh" (the only method diffxpy itself ever requests, det.py:115).
    :param alpha: unused by fdr_bh q-values, kept for signature compatibility.
    """
    if method != "fdr_bh":
        raise ValueError("multiple testing correction %r is not available (only 'fdr_bh')" % method)
    dev = cuda_device()
    p = torch.from_numpy(np.ascontiguousarray(np.asarray(pvals, dtype=np.float64))).to(dev)
    return bh_device(p).cpu().numpy()
