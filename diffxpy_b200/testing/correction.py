"""
Multiple-testing correction (/root/reference/diffxpy/testing/correction.py:5-24).  The reference calls
statsmodels.stats.multitest.multipletests(method="fdr_bh") on the non-NaN p-values; the same arithmetic
(p sorted ascending, q = p / (i/n), reverse running minimum, clip at 1, unsort) runs on the GPU here.
"""
import numpy as np
import torch

from ..device import cuda_device


def bh_device(p):
    """Benjamini-Hochberg on a float64 CUDA vector; NaNs are excluded from n and stay NaN."""
    q = torch.full_like(p, float("nan"))
    ok = ~torch.isnan(p)
    pv = p[ok]
    n = pv.shape[0]
    if n == 0:
        return q
    ps, order = torch.sort(pv, stable=True)
    ecdf = torch.arange(1, n + 1, dtype=torch.float64, device=p.device) / float(n)
    qs = ps / ecdf
    qs = torch.flip(torch.cummin(torch.flip(qs, [0]), 0).values, [0])
    qs = torch.clamp(qs, max=1.0)
    out = torch.empty_like(qs)
    out[order] = qs
    q[ok] = out
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
