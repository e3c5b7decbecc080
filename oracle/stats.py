"""
CPU oracle for the test statistics.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, function by function, /root/reference/diffxpy/stats/stats.py and the model-free
result constructors of /root/reference/diffxpy/testing/det.py.  PINNED: checked against vectors
produced by the reference file itself and scipy (tests/golden/make_golden.py).
"""
import sys
import numpy as np
import scipy.sparse
import scipy.special
import scipy.stats

TINY = np.nextafter(0.0, np.inf)
DMAX = np.nextafter(np.inf, 0.0)


def wald_test(theta_mle, theta_sd, theta0=0.0):
    """stats.py:181-218 -- p = 2 (1 - Phi(|theta - theta0| / max(sd, tiny)))."""
    theta_mle = np.asarray(theta_mle, dtype=np.float64)
    theta_sd = np.maximum(np.asarray(theta_sd, dtype=np.float64), TINY)
    z = np.abs((theta_mle - theta0) / theta_sd)
    return 2.0 * (1.0 - scipy.stats.norm.cdf(z))


def wald_test_chisq(theta_mle, theta_covar, theta0=0.0):
    """stats.py:221-282 -- w = d^T Sigma^-1 d per gene, NaN where cond(Sigma) >= 1/eps,
    p = 1 - chi2_k.cdf(w).  theta_mle (par x genes), theta_covar (genes x par x par)."""
    theta_mle = np.asarray(theta_mle, dtype=np.float64)
    theta_covar = np.asarray(theta_covar, dtype=np.float64)
    d = theta_mle - theta0
    g = theta_covar.shape[0]
    w = np.full(g, np.nan)
    with np.errstate(all="ignore"):
        for i in range(g):
            if not np.all(np.isfinite(theta_covar[i])):
                continue
            if np.linalg.cond(theta_covar[i]) < 1.0 / sys.float_info.epsilon:
                w[i] = d[:, i] @ np.linalg.inv(theta_covar[i]) @ d[:, i]
    return 1.0 - scipy.stats.chi2(theta_mle.shape[0]).cdf(w)


def likelihood_ratio_test(ll_full, ll_reduced, df_full, df_reduced):
    """stats.py:9-39 -- p = 1 - chi2_{ddf}.cdf(2 (ll_full - ll_reduced))."""
    delta_df = df_full - df_reduced
    delta_dev = 2.0 * (np.asarray(ll_full, dtype=np.float64) - np.asarray(ll_reduced, dtype=np.float64))
    return 1.0 - scipy.stats.chi2(delta_df).cdf(delta_dev)


def two_coef_z_test(theta_mle0, theta_mle1, theta_sd0, theta_sd1):
    """stats.py:285-326."""
    div = np.sqrt(np.maximum(np.square(theta_sd0) + np.square(theta_sd1), TINY))
    z = np.abs(theta_mle0 - theta_mle1) / div
    return 2.0 * (1.0 - scipy.stats.norm.cdf(z))


def t_test_moments(mu0, mu1, var0, var1, n0, n1):
    """stats.py:116-178 -- Welch t from moments with the reference's clipping."""
    mu0, mu1, var0, var1 = (np.asarray(v, dtype=np.float64) for v in (mu0, mu1, var0, var1))
    s_delta = np.clip(np.sqrt(var0 / n0 + var1 / n1), TINY, DMAX)
    t = np.abs(mu0 - mu1) / s_delta
    divisor = np.clip(np.square(var0 / n0) / (n0 - 1) + np.square(var1 / n1) / (n1 - 1), TINY, DMAX)
    df = np.clip(np.square(var0 / n0 + var1 / n1) / divisor, TINY, DMAX)
    return 2.0 * scipy.stats.t.sf(t, df)


def mann_whitney_u(x, y):
    """scipy.stats.mannwhitneyu(x, y, use_continuity=True, alternative='two-sided') restated for the
    asymptotic method with integer-exact intermediates (scipy/stats/_mannwhitneyu.py).
    Returns (U1, 2*R1 as int, tie term sum(t^3 - t) as int, pvalue)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    n1, n2 = len(x), len(y)
    n = n1 + n2
    allv = np.concatenate([x, y])
    uniq, inv, cnt = np.unique(allv, return_inverse=True, return_counts=True)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    twice_rank = 2 * start + cnt + 1          # 2 * average rank of each tie block (exact integer)
    r1x2 = int(np.sum(twice_rank[inv[:n1]].astype(object)))
    tie = int(np.sum((cnt.astype(object) ** 3 - cnt.astype(object))))
    u1 = r1x2 / 2.0 - n1 * (n1 + 1) / 2.0
    mu = n1 * n2 / 2.0
    s = np.sqrt(n1 * n2 / 12.0 * ((n + 1) - tie / (n * (n - 1.0))))
    num = u1 - mu
    num = num - 0.5 * np.sign(num)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.float64(num) / s
    p = float(np.clip(2.0 * scipy.special.ndtr(-np.abs(z)), 0.0, 1.0))
    return u1, r1x2, tie, p


def mann_whitney_u_test(x0, x1):
    """stats.py:42-75 -- per-gene loop over scipy.stats.mannwhitneyu (the reference's own call)."""
    if scipy.sparse.issparse(x0):
        x0 = np.asarray(x0.todense())
        x1 = np.asarray(x1.todense())
    return np.asarray([
        scipy.stats.mannwhitneyu(x0[:, i], x1[:, i], use_continuity=True, alternative="two-sided").pvalue
        for i in range(x0.shape[1])
    ])


def bh_correct(pvals):
    """testing/correction.py:5-24 with statsmodels' fdr_bh arithmetic (multitest.py: p sorted ascending,
    q = p / (i/n), reverse running minimum, clip at 1, unsort); NaN p-values are excluded from n."""
    pvals = np.asarray(pvals, dtype=np.float64)
    q = np.full(pvals.shape[0], np.nan)
    ok = ~np.isnan(pvals)
    p = pvals[ok]
    n = p.shape[0]
    if n == 0:
        return q
    order = np.argsort(p, kind="stable")
    ps = p[order]
    ecdf = np.arange(1, n + 1) / float(n)
    qs = ps / ecdf
    qs = np.minimum.accumulate(qs[::-1])[::-1]
    qs[qs > 1] = 1
    out = np.empty(n)
    out[order] = qs
    q[ok] = out
    return q


def split_x(data, grouping):
    """testing/utils.py:112-117 -- groups in order of first appearance."""
    import pandas as pd
    grouping = np.asarray(grouping)
    groups = pd.Series(grouping).unique()
    x0 = data[np.where(grouping == groups[0])[0]]
    x1 = data[np.where(grouping == groups[1])[0]]
    return x0, x1


def two_group_test(data, grouping, test, is_logged=False, is_sig_zerovar=True):
    """det.py:1545-1620 (t-test) and det.py:1672-1743 (rank test) constructors.
    Returns dict(pval, mean, logfc, zero_variance, mean_x0, mean_x1, var_x0, var_x1)."""
    x0, x1 = split_x(data, grouping)
    n0, n1 = x0.shape[0], x1.shape[0]
    mean_x0 = np.asarray(x0.mean(axis=0)).flatten().astype(float)
    mean_x1 = np.asarray(x1.mean(axis=0)).flatten().astype(float)
    mean = np.average(np.vstack([mean_x0, mean_x1]), weights=np.array([n0 / (n0 + n1), n1 / (n0 + n1)]), axis=0)
    if scipy.sparse.issparse(x0):
        var_x0 = np.asarray(x0.power(2).mean(axis=0)).flatten().astype(float) - np.square(mean_x0)
        var_x1 = np.asarray(x1.power(2).mean(axis=0)).flatten().astype(float) - np.square(mean_x1)
    else:
        var_x0 = np.asarray(np.var(x0, axis=0)).flatten().astype(float)
        var_x1 = np.asarray(np.var(x1, axis=0)).flatten().astype(float)
    var_geq_zero = np.logical_or(var_x0 > 0, var_x1 > 0)
    idx_run = np.where(np.logical_and(mean != 0, var_geq_zero))[0]
    pval = np.zeros(data.shape[1]) + np.nan
    if test == "t":
        pval[idx_run] = t_test_moments(mean_x0[idx_run], mean_x1[idx_run], var_x0[idx_run], var_x1[idx_run], n0, n1)
    elif test == "rank":
        pval[idx_run] = mann_whitney_u_test(x0[:, idx_run], x1[:, idx_run])
    else:
        raise ValueError(test)
    pval[np.where(np.logical_and(np.logical_and(mean_x0 == mean_x1, mean > 0), ~var_geq_zero))[0]] = 1.0
    if is_sig_zerovar:
        pval[np.where(np.logical_and(mean_x0 != mean_x1, ~var_geq_zero))[0]] = 0.0
    if is_logged:
        logfc = mean_x1 - mean_x0
    else:
        logfc = np.log(np.maximum(mean_x1, TINY)) - np.log(np.maximum(mean_x0, TINY))
    return dict(pval=pval, mean=mean, logfc=logfc, zero_variance=~var_geq_zero,
                mean_x0=mean_x0, mean_x1=mean_x1, var_x0=var_x0, var_x1=var_x1, n0=n0, n1=n1)
