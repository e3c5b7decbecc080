"""
Multiple-testing correction (/root/reference/diffxpy/testing/correction.py:5-24).  The reference calls
statsmodels.stats.multitest.multipletests(method="fdr_bh") on the non-NaN p-values; the same arithmetic
(p sorted ascending, q = p / (i/n), reverse running minimum, clip at 1, unsort) runs on the GPU here.
"""
import numpy as np
import torch

from .. import _lib
from ..device import cuda_device, ptr, stream_ptr

BH_KERNEL_MAX = 65536     # above this the all-pairs kernel loses to a device sort


def bh_device(p):
    """Benjamini-Hochberg on a float64 CUDA vector; NaNs are excluded from n and stay NaN."""
    if p.shape[0] <= BH_KERNEL_MAX:
        p = p.contiguous()
        q = torch.empty_like(p)
        work = torch.empty((p.shape[0] * 3 + 1) // 2, dtype=torch.float64, device=p.device)   # n doubles + n int32
        _lib.call("dx_bh_qvalues", p.shape[0], ptr(p), ptr(q), ptr(work), stream_ptr())
        return q
    # sort path, no host synchronisation: NaNs sort last and are masked out of the running minimum
    g = p.shape[0]
    ps, order = torch.sort(p, stable=True)
    bad = torch.isnan(ps)
    n = (g - bad.sum()).to(torch.float64)                       # 0-d tensor: true division like numpy's
    ecdf = torch.arange(1, g + 1, dtype=torch.float64, device=p.device) / n
    qs = torch.where(bad, torch.full_like(ps, float("inf")), ps / ecdf)
    qs = torch.flip(torch.cummin(torch.flip(qs, [0]), 0).values, [0])
    qs = torch.where(bad, torch.full_like(ps, float("nan")), torch.clamp(qs, max=1.0))
    q = torch.empty_like(qs)
    q[order] = qs
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
