"""
de.stats.* -- the reference's vectorised test statistics (/root/reference/diffxpy/stats/stats.py), same names,
arguments and error behaviour, evaluated by the CUDA kernels of libdxb200.so (K_stats / K_twogroup).
Inputs and outputs are numpy arrays like in the reference; there is no CPU fallback.
"""
from typing import Union

import numpy as np
import scipy.sparse
import scipy.stats
import torch

from .. import _lib
from ..device import CountShard, ptr, stream_ptr, cuda_device


def _dev(x, dev):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)


def likelihood_ratio_test(ll_full: np.ndarray, ll_reduced: np.ndarray, df_full: int, df_reduced: int):
    """stats.py:9-39: p = 1 - chi2(df_full - df_reduced).cdf(2 (ll_full - ll_reduced))."""
    if ll_full.shape[0] != ll_full.shape[0]:
        raise ValueError('stats.likelihood_ratio_test(): ll_full and ll_red have to contain the same number of genes')
    return likelihood_ratio_test_full(ll_full, ll_reduced, df_full, df_reduced)[0]


def likelihood_ratio_test_full(ll_full, ll_reduced, df_full, df_reduced):
    """(pval in the reference's 1 - cdf form, accurate log pval)."""
    dev = cuda_device()
    a, b = _dev(ll_full, dev), _dev(ll_reduced, dev)
    p, lp = torch.empty_like(a), torch.empty_like(a)
    _lib.call("dx_lrt", a.shape[0], ptr(a), ptr(b), int(df_full - df_reduced), ptr(p), ptr(lp), stream_ptr())
    return p.cpu().numpy(), lp.cpu().numpy()


def wald_test(theta_mle: np.ndarray, theta_sd: np.ndarray, theta0: Union[int, np.ndarray] = 0):
    """stats.py:181-218: two-sided normal test of (theta_mle - theta0) / max(theta_sd, tiny)."""
    return wald_test_full(theta_mle, theta_sd, theta0)[0]


def wald_test_full(theta_mle, theta_sd, theta0=0):
    theta_mle = np.asarray(theta_mle, dtype=np.float64)
    theta_sd = np.asarray(theta_sd, dtype=np.float64)
    if np.size(theta0) == 1:
        theta0 = np.broadcast_to(theta0, theta_mle.shape)
    if theta_mle.shape[0] != theta_sd.shape[0]:
        raise ValueError('stats.wald_test(): theta_mle and theta_sd have to contain the same number of entries')
    if theta0.shape[0] > 1:
        if theta_mle.shape[0] != theta0.shape[0]:
            raise ValueError('stats.wald_test(): theta_mle and theta0 have to contain the same number of entries')
    dev = cuda_device()
    t, s = _dev(theta_mle - theta0, dev), _dev(theta_sd, dev)
    p, lp = torch.empty_like(t), torch.empty_like(t)
    _lib.call("dx_wald_test", t.shape[0], ptr(t), ptr(s), ptr(p), ptr(lp), stream_ptr())
    return p.cpu().numpy(), lp.cpu().numpy()


def wald_test_chisq(theta_mle: np.ndarray, theta_covar: np.ndarray, theta0: Union[int, np.ndarray] = 0):
    """stats.py:221-282: multi-coefficient Wald test, NaN where the covariance is numerically singular."""
    return wald_test_chisq_full(theta_mle, theta_covar, theta0)[0]


def wald_test_chisq_full(theta_mle, theta_covar, theta0=0):
    theta_mle = np.asarray(theta_mle, dtype=np.float64)
    theta_covar = np.asarray(theta_covar, dtype=np.float64)
    if np.size(theta0) == 1:
        theta0 = np.broadcast_to(theta0, theta_mle.shape)
    if theta_mle.shape[0] != theta_covar.shape[1]:
        raise ValueError('stats.wald_test(): theta_mle and theta_covar have to contain the same number of parameters')
    if len(theta_covar.shape) != 3:
        raise ValueError('stats.wald_test(): theta_covar should have 3 dimensions but has %i' % len(theta_covar.shape))
    if theta_mle.shape[1] != theta_covar.shape[0]:
        raise ValueError('stats.wald_test(): theta_mle and theta_covar have to contain the same number of genes')
    if theta_covar.shape[1] != theta_covar.shape[2]:
        raise ValueError('stats.wald_test(): the first two dimensions of theta_covar have to be of the same size')
    if theta0.shape[0] > 1:
        if theta_mle.shape[0] != theta0.shape[0]:
            raise ValueError('stats.wald_test(): theta_mle and theta0 have to contain the same number of entries')
        if theta_mle.shape[1] != theta0.shape[1]:
            raise ValueError('stats.wald_test(): theta_mle and theta0 have to contain the same number of entries')
    dev = cuda_device()
    d, cov = _dev(theta_mle - theta0, dev), _dev(theta_covar, dev)
    g = cov.shape[0]
    stat = torch.empty(g, dtype=torch.float64, device=dev)
    p, lp = torch.empty_like(stat), torch.empty_like(stat)
    _lib.call("dx_wald_chisq", g, int(theta_mle.shape[0]), ptr(d), ptr(cov), ptr(stat), ptr(p), ptr(lp), stream_ptr())
    p = p.cpu().numpy()
    n_bad = int(np.isnan(stat.cpu().numpy()).sum())
    if n_bad > 0:
        print("caught %i linalg singular matrix errors" % n_bad)
    return p, lp.cpu().numpy()


def two_coef_z_test(theta_mle0: np.ndarray, theta_mle1: np.ndarray, theta_sd0: np.ndarray, theta_sd1: np.ndarray):
    """stats.py:285-326."""
    if theta_mle0.shape[0] != theta_mle1.shape[0]:
        raise ValueError('stats.two_coef_z_test(): theta_mle0 and theta_mle1 have to contain the same number of entries')
    if theta_sd0.shape[0] != theta_sd1.shape[0]:
        raise ValueError('stats.two_coef_z_test(): theta_sd0 and theta_sd1 have to contain the same number of entries')
    if theta_mle0.shape[0] != theta_sd0.shape[0]:
        raise ValueError('stats.two_coef_z_test(): theta_mle0 and theta_sd0 have to contain the same number of entries')
    dev = cuda_device()
    a, b, c, d = (_dev(v, dev) for v in (theta_mle0, theta_mle1, theta_sd0, theta_sd1))
    p = torch.empty_like(a)
    _lib.call("dx_two_coef_z", a.shape[0], ptr(a), ptr(b), ptr(c), ptr(d), ptr(p), stream_ptr())
    return p.cpu().numpy()


def t_test_moments(mu0: np.ndarray, mu1: np.ndarray, var0: np.ndarray, var1: np.ndarray, n0: int, n1: int):
    """stats.py:116-178: Welch's t-test from group moments."""
    if len(mu0) != len(mu1):
        raise ValueError('stats.t_test_moments(): mu and mu1 have to contain the same number of entries')
    if len(var0) != len(var1):
        raise ValueError('stats.t_test_moments(): mu and mu1 have to contain the same number of entries')
    dev = cuda_device()
    a, b, c, d = (_dev(v, dev) for v in (mu0, mu1, var0, var1))
    p = torch.empty_like(a)
    _lib.call("dx_t_test_moments", a.shape[0], ptr(a), ptr(b), ptr(c), ptr(d), float(n0), float(n1), ptr(p), stream_ptr())
    return p.cpu().numpy()


def t_test_raw(x0, x1):
    """stats.py:78-113: Welch's t-test on raw data (observations x genes)."""
    if x0.shape[1] != x1.shape[1]:
        raise ValueError('stats.t_test_raw(): x0 and x1 have to contain the same number of genes')
    res = two_group_raw(x0, x1, do_rank=False)
    return res["out"][:, _lib.TG_PVAL_T]


def mann_whitney_u_test(x0: Union[np.ndarray, scipy.sparse.csr_matrix], x1: Union[np.ndarray, scipy.sparse.csr_matrix]):
    """stats.py:42-75: Mann-Whitney U / Wilcoxon rank-sum test per gene (scipy.stats.mannwhitneyu asymptotic
    method, continuity and tie corrected, two-sided)."""
    if np.any(np.ndim(x0) != np.ndim(x1)):
        raise ValueError('number of dimensions is not allowed to differ between x0 and x1')
    if np.ndim(x0) == 1:
        x0 = x0.reshape([x0.shape[0], 1])
        x1 = x1.reshape([x1.shape[0], 1])
    if np.any(x0.shape[1] != x1.shape[1]):
        raise ValueError('the first axis (number of tests) is not allowed to differ between x0 and x1')
    res = two_group_raw(x0, x1, do_rank=True)
    if np.any(res["flags"] & _lib.FLAG_RANK_OVERFLOW):
        raise _lib.DxError("rank test: a gene has more non-count values than the rank workspace holds (stats.RANK_WORK_MAX_BYTES)")
    return res["out"][:, _lib.TG_PVAL_RANK]


def two_group_raw(x0, x1, do_rank=True):
    """Run K_ingest (2 groups) + K_twogroup on the stacked samples x0 (first sample) and x1."""
    if scipy.sparse.issparse(x0):
        x = scipy.sparse.vstack([x0, x1]).tocsc()
    else:
        x = np.concatenate([np.asarray(x0), np.asarray(x1)], axis=0)
    grouping = np.concatenate([np.zeros(x0.shape[0], dtype=np.uint8), np.ones(x1.shape[0], dtype=np.uint8)])
    return two_group_device(CountShard.from_host(x), grouping, x0.shape[0], x1.shape[0], do_rank)


def two_group_device(shard, cell_group, n0, n1, do_rank=True):
    dev = shard.device
    with torch.cuda.device(dev):
        cg = torch.from_numpy(np.ascontiguousarray(cell_group, dtype=np.uint8)).to(dev)
        suff = shard.group_suffstats(cg, 2)
        g = shard.n_genes
        out = torch.empty((g, _lib.TG_NCOL), dtype=torch.float64, device=dev)
        u1x2 = torch.empty(g, dtype=torch.int64, device=dev)
        tie = torch.empty(g, dtype=torch.int64, device=dev)
        flags = torch.empty(g, dtype=torch.int32, device=dev)
        _lib.call("dx_two_group_tests", g, float(n0), float(n1), ptr(suff.tail), ptr(suff.ovf_count), ptr(shard.indptr),
                  ptr(suff.ovf_val), ptr(suff.ovf_group), int(bool(do_rank)), ptr(out), ptr(u1x2), ptr(tie), ptr(flags),
                  stream_ptr())
        if do_rank:
            _rank_long_lists(suff, 0, 1, n0, n1, out, u1x2, tie, flags)
        return dict(out=out.cpu().numpy(), u1x2=u1x2.cpu().numpy(), tie=tie.cpu().numpy(), flags=flags.cpu().numpy())


RANK_WORK_MAX_BYTES = 4 << 30      # device workspace the rank test of long non-count lists may take


def _rank_long_lists(suff, group_a, group_b, n_a, n_b, out, u1x2, tie, flags):
    """Genes K_twogroup flagged FLAG_RANK_OVERFLOW (more non-count values than its on-chip sort holds: normalised /
    log-transformed data, highly expressed genes) are ranked by dx_group_pair_tests_big in a device workspace."""
    flagged = (flags & _lib.FLAG_RANK_OVERFLOW) != 0
    n_flag = int(flagged.sum().item())
    if n_flag == 0:
        return
    shard = suff.shard
    longest = int(suff.ovf_count[flagged].max().item()) + _lib.DX_HIST_BINS + 1
    slice_elems = 8192
    while slice_elems < longest:
        slice_elems *= 2
    sm = torch.cuda.get_device_properties(shard.device).multi_processor_count
    n_cta = max(1, min(n_flag, 2 * sm, RANK_WORK_MAX_BYTES // (8 * slice_elems)))
    work = torch.empty(n_cta * slice_elems, dtype=torch.int64, device=shard.device)
    _lib.call("dx_group_pair_tests_big", shard.n_genes, int(suff.n_groups), int(group_a), int(group_b), float(n_a),
              float(n_b), ptr(suff.tail), ptr(suff.ovf_count), ptr(shard.indptr), ptr(suff.ovf_val), ptr(suff.ovf_group),
              ptr(out), ptr(u1x2), ptr(tie), ptr(flags), ptr(work), work.numel() * 8, slice_elems * 8, stream_ptr())


def group_pair_device(suff, group_a, group_b, n_a, n_b, do_rank=True):
    """K_twogroup between two of the samples of ONE K_ingest pass (``suff``: SuffStats over n_groups samples):
    sample 0 = group_a, sample 1 = group_b or, for group_b < 0, the union of all other groups."""
    shard = suff.shard
    if suff.ovf_packed is not None:
        raise ValueError("group_pair_device: these statistics were packed for the NB fit (SuffStats.pack_overflow)")
    dev = shard.device
    with torch.cuda.device(dev):
        g = shard.n_genes
        out = torch.empty((g, _lib.TG_NCOL), dtype=torch.float64, device=dev)
        u1x2 = torch.empty(g, dtype=torch.int64, device=dev)
        tie = torch.empty(g, dtype=torch.int64, device=dev)
        flags = torch.empty(g, dtype=torch.int32, device=dev)
        _lib.call("dx_group_pair_tests", g, int(suff.n_groups), int(group_a), int(group_b), float(n_a), float(n_b),
                  ptr(suff.tail), ptr(suff.ovf_count), ptr(shard.indptr), ptr(suff.ovf_val), ptr(suff.ovf_group),
                  int(bool(do_rank)), ptr(out), ptr(u1x2), ptr(tie), ptr(flags), stream_ptr())
        if do_rank:
            _rank_long_lists(suff, group_a, group_b, n_a, n_b, out, u1x2, tie, flags)
        return dict(out=out.cpu().numpy(), u1x2=u1x2.cpu().numpy(), tie=tie.cpu().numpy(), flags=flags.cpu().numpy())


def hypergeom_test(intersections: np.ndarray, enquiry: int, references: np.ndarray, background: int) -> np.ndarray:
    """stats.py:329-355: hypergeometric (gene-set over-representation) test; a handful of scalars per gene set, evaluated
    on the host with the arithmetic of the reference (1 - cdf(x - 1) of scipy.stats.hypergeom)."""
    return np.array([1 - scipy.stats.hypergeom(M=background, n=references[i], N=enquiry).cdf(x - 1)
                     for i, x in enumerate(intersections)])
