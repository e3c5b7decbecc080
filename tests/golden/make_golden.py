"""
Generates the golden fixtures under tests/golden/.  Run in the build container only
(it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

* stats_golden.npz   -- outputs of the REFERENCE's own diffxpy/stats/stats.py (loaded straight from
                        /root/reference; it needs only numpy + scipy) and of scipy on seeded inputs.
                        These pin oracle/stats.py and the CUDA statistics kernels.
* nbglm_cfg1.npz     -- BASELINE.json configs[0] (2000 cells x 100 genes, '~ 1 + condition') fitted by
                        oracle/nb_glm.py.  ORACLE = RESTATEMENT, NOT batchglm: the GLM arithmetic of
                        the reference lives in the absent third-party package batchglm (parity unpinned).
* nbglm_batch.npz    -- 1500 cells x 60 genes, '~ 1 + condition + batch' (4 batches), low counts,
                        same caveat.
"""
import importlib.util
import os
import sys

import numpy as np
import scipy.sparse
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF_STATS = "/root/reference/diffxpy/stats/stats.py"


def load_ref_stats():
    spec = importlib.util.spec_from_file_location("ref_stats", REF_STATS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_stats():
    ref = load_ref_stats()
    rng = np.random.default_rng(1)
    out = {}
    # --- wald_test (stats.py:181-218): include saturation (|z| > 8.3), tiny sd, zero coefficient
    g = 400
    theta = rng.normal(0, 1.0, g)
    sd = np.exp(rng.normal(-1.5, 1.0, g))
    theta[:4] = [0.0, 5.0, -40.0, 1e-12]
    sd[:4] = [0.3, 0.1, 0.5, 0.0]
    out["wald_theta"], out["wald_sd"] = theta.copy(), sd.copy()
    out["wald_p"] = ref.wald_test(theta_mle=theta.copy(), theta_sd=sd.copy(), theta0=0)
    # NaN standard deviations (singular Fisher information) and NaN coefficients stay NaN in the reference: the
    # np.nextafter(..., where=sd < tiny) floor does not touch them
    th_n = np.array([0.7, -2.0, np.nan, 1.5, 0.0])
    sd_n = np.array([np.nan, 0.4, 0.3, np.nan, 0.0])
    out["waldnan_theta"], out["waldnan_sd"] = th_n.copy(), sd_n.copy()
    with np.errstate(all="ignore"):
        out["waldnan_p"] = ref.wald_test(theta_mle=th_n.copy(), theta_sd=sd_n.copy(), theta0=0)
    # --- wald_test_chisq (stats.py:221-282): k = 3 coefficients, one singular covariance
    k = 3
    th = rng.normal(0, 0.4, (k, g))
    l = rng.normal(0, 0.2, (g, k, k))
    cov = np.einsum('gij,gkj->gik', l, l) + 0.01 * np.eye(k)[None]
    cov[7] = np.ones((k, k))          # singular -> NaN
    out["chisq_theta"], out["chisq_cov"] = th.copy(), cov.copy()
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        out["chisq_p"] = ref.wald_test_chisq(theta_mle=th.copy(), theta_covar=cov.copy(), theta0=0)
    finally:
        sys.stdout.close()
        sys.stdout = _stdout
    # --- likelihood_ratio_test (stats.py:9-39)
    ll_red = -rng.uniform(1e3, 1e5, g)
    dev = rng.chisquare(2, g)
    dev[:3] = [0.0, 200.0, -1e-3]
    ll_full = ll_red + dev / 2
    out["lrt_ll_full"], out["lrt_ll_red"] = ll_full, ll_red
    out["lrt_df"] = np.array([9, 7])
    out["lrt_p"] = ref.likelihood_ratio_test(ll_full=ll_full, ll_reduced=ll_red, df_full=9, df_reduced=7)
    # --- t_test_moments (stats.py:116-178)
    mu0 = rng.gamma(2.0, 1.0, g)
    mu1 = mu0 * np.exp(rng.normal(0, 0.2, g))
    var0 = mu0 * rng.uniform(0.5, 3.0, g)
    var1 = mu1 * rng.uniform(0.5, 3.0, g)
    var0[:2] = [0.0, 1e-300]
    out["tt_mu0"], out["tt_mu1"], out["tt_var0"], out["tt_var1"] = mu0, mu1, var0, var1
    out["tt_n"] = np.array([731, 1269])
    out["tt_p"] = ref.t_test_moments(mu0=mu0, mu1=mu1, var0=var0, var1=var1, n0=731, n1=1269)
    # --- two_coef_z_test (stats.py:285-326)
    out["z_p"] = ref.two_coef_z_test(theta_mle0=theta.copy(), theta_mle1=theta[::-1].copy(),
                                     theta_sd0=sd.copy() + 0.1, theta_sd1=sd[::-1].copy() + 0.1)
    # --- mann_whitney_u_test (stats.py:42-75) on tie-heavy count data, dense and CSR, plus edge genes
    n, gg = 600, 40
    lam = np.exp(rng.uniform(np.log(0.05), np.log(8.0), gg))
    x = rng.poisson(rng.gamma(1.5, lam / 1.5, size=(n, gg))).astype(np.float64)
    x[:, 1] = 0.0
    x[5, 1] = 3.0                       # single non-zero
    x[:, 2] = np.round(rng.normal(0, 1.5, n))   # negative values
    x[:, 3] = rng.normal(0, 1, n)        # continuous, no ties
    grouping = rng.integers(0, 2, n)
    grouping[0] = 1                      # first-appearance order: group "1" is x0
    x0 = x[grouping == 1]
    x1 = x[grouping == 0]
    out["mwu_x"], out["mwu_grouping"] = x, grouping
    out["mwu_p_dense"] = ref.mann_whitney_u_test(x0, x1)
    out["mwu_p_csr"] = ref.mann_whitney_u_test(scipy.sparse.csr_matrix(x0), scipy.sparse.csr_matrix(x1))
    out["mwu_u1"] = np.array([scipy.stats.mannwhitneyu(x0[:, i], x1[:, i], use_continuity=True,
                                                        alternative="two-sided").statistic for i in range(gg)])
    # --- BH: statsmodels is absent; scipy's false_discovery_control(method='bh') is the installed cross-check
    p = rng.uniform(0, 1, 300) ** 2
    p[:5] = [0.0, 1.0, 1e-300, 0.5, 0.5]
    out["bh_p"] = p
    out["bh_q"] = scipy.stats.false_discovery_control(p, method="bh")
    np.savez_compressed(os.path.join(HERE, "stats_golden.npz"), **out)
    print("stats_golden.npz:", {k_: v.shape for k_, v in out.items()})


def simulate_nb(rng, n, g, log_mu_range, r_range, cond, batch=None, frac_de=0.5, sd_cond=0.5, sd_batch=0.2,
                size_factors=None):
    log_mu = rng.uniform(np.log(log_mu_range[0]), np.log(log_mu_range[1]), g)
    r = rng.uniform(r_range[0], r_range[1], g)
    beta_c = np.where(np.arange(g) >= int(g * (1 - frac_de)), rng.normal(0, sd_cond, g), 0.0)
    eta = log_mu[None, :] + cond[:, None] * beta_c[None, :]
    if batch is not None:
        nb_ = int(batch.max()) + 1
        beta_b = rng.normal(0, sd_batch, (nb_, g))
        beta_b[0] = 0
        eta = eta + beta_b[batch]
    if size_factors is not None:
        eta = eta + np.log(size_factors)[:, None]
    mu = np.exp(eta)
    y = rng.poisson(rng.gamma(r[None, :], mu / r[None, :])).astype(np.float64)
    return y, log_mu, r, beta_c


def make_nbglm():
    import oracle.nb_glm as nb
    # ---- configs[0]: 2000 x 100, ~1 + condition
    rng = np.random.default_rng(1)
    n, g = 2000, 100
    cond = (np.arange(n) % 2).astype(np.float64)
    y, log_mu, r, beta_c = simulate_nb(rng, n, g, (100, 1000), (1, 2), cond)
    xd = np.stack([np.ones(n), cond], axis=1)
    zd = np.ones((n, 1))
    out = dict(x=y.astype(np.float32), design_loc=xd, design_scale=zd)
    for mode in ("brent", "newton"):
        fit = nb.fit_nb(y, xd, zd, method_b=mode)
        for key in ("a_var", "b_var", "fisher_inv", "log_likelihood", "jacobian", "niter", "error_codes"):
            out["%s_%s" % (mode, key)] = fit[key]
    np.savez_compressed(os.path.join(HERE, "nbglm_cfg1.npz"), **out)
    print("nbglm_cfg1.npz written; niter", np.unique(out["newton_niter"], return_counts=True))
    # ---- low counts, batch design
    rng = np.random.default_rng(2)
    n, g = 1500, 60
    cond = rng.integers(0, 2, n).astype(np.float64)
    batch = rng.integers(0, 4, n)
    y, log_mu, r, beta_c = simulate_nb(rng, n, g, (0.05, 5.0), (0.5, 4.0), cond, batch)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 4)[None, :])], axis=1).astype(float)
    zd = np.ones((n, 1))
    out = dict(x=y.astype(np.float32), design_loc=xd, design_scale=zd, cond=cond, batch=batch)
    for mode in ("brent", "newton"):
        fit = nb.fit_nb(y, xd, zd, method_b=mode)
        for key in ("a_var", "b_var", "fisher_inv", "log_likelihood", "jacobian", "niter", "error_codes"):
            out["%s_%s" % (mode, key)] = fit[key]
    np.savez_compressed(os.path.join(HERE, "nbglm_batch.npz"), **out)
    print("nbglm_batch.npz written; niter", np.unique(out["newton_niter"], return_counts=True),
          "flags", np.unique(out["newton_error_codes"], return_counts=True))


if __name__ == "__main__":
    make_stats()
    if "--stats-only" not in sys.argv:
        make_nbglm()
