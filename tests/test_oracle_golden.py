"""CPU: the oracle against the golden vectors produced by the reference's own stats.py / scipy
(tests/golden/make_golden.py), plus internal consistency of the NB GLM restatement."""
import os

import numpy as np
import pytest
import scipy.sparse
import scipy.stats

import oracle.nb_glm as onb
import oracle.stats as ost


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stats_golden.npz"))


def test_wald_test_matches_reference(gold):
    p = ost.wald_test(gold["wald_theta"], gold["wald_sd"])
    np.testing.assert_allclose(p, gold["wald_p"], rtol=0, atol=0)      # same scipy call -> bit identical


def test_wald_chisq_matches_reference(gold):
    p = ost.wald_test_chisq(gold["chisq_theta"], gold["chisq_cov"])
    assert np.isnan(p[7]) and np.isnan(gold["chisq_p"][7])
    ok = ~np.isnan(gold["chisq_p"])
    np.testing.assert_allclose(p[ok], gold["chisq_p"][ok], rtol=1e-12, atol=1e-15)


def test_lrt_matches_reference(gold):
    p = ost.likelihood_ratio_test(gold["lrt_ll_full"], gold["lrt_ll_red"], int(gold["lrt_df"][0]), int(gold["lrt_df"][1]))
    np.testing.assert_allclose(p, gold["lrt_p"], rtol=0, atol=0)


def test_wald_nan_sd_stays_nan(gold):
    """A NaN standard deviation (singular Fisher information) gives p = NaN in the reference (stats.py:215 floors only
    sd < tiny), and such genes are left out of the BH correction (correction.py:15-22)."""
    with np.errstate(all="ignore"):
        p = ost.wald_test(gold["waldnan_theta"], gold["waldnan_sd"])
    np.testing.assert_array_equal(p, gold["waldnan_p"])
    assert np.isnan(p[0]) and np.isnan(p[2]) and np.isnan(p[3]) and p[4] == 1.0
    q = ost.bh_correct(p)
    assert np.isnan(q[0]) and np.isfinite(q[1])
    np.testing.assert_allclose(q[[1, 4]], ost.bh_correct(p[[1, 4]]), rtol=0, atol=0)


def test_t_test_moments_matches_reference(gold):
    p = ost.t_test_moments(gold["tt_mu0"], gold["tt_mu1"], gold["tt_var0"], gold["tt_var1"],
                           int(gold["tt_n"][0]), int(gold["tt_n"][1]))
    np.testing.assert_allclose(p, gold["tt_p"], rtol=0, atol=0)


def test_two_coef_z_matches_reference(gold):
    th, sd = gold["wald_theta"], gold["wald_sd"]
    p = ost.two_coef_z_test(th, th[::-1], sd + 0.1, sd[::-1] + 0.1)
    np.testing.assert_allclose(p, gold["z_p"], rtol=1e-14, atol=0)


def test_mann_whitney_matches_reference_and_is_integer_exact(gold):
    x, grouping = gold["mwu_x"], gold["mwu_grouping"]
    x0, x1 = x[grouping == 1], x[grouping == 0]
    np.testing.assert_array_equal(gold["mwu_p_dense"], gold["mwu_p_csr"])
    for i in range(x.shape[1]):
        u1, r1x2, tie, p = ost.mann_whitney_u(x0[:, i], x1[:, i])
        assert u1 == gold["mwu_u1"][i]                       # bit exact U statistic
        if np.isnan(gold["mwu_p_dense"][i]):
            assert np.isnan(p)
        else:
            np.testing.assert_allclose(p, gold["mwu_p_dense"][i], rtol=1e-13, atol=0)


def test_bh_matches_scipy(gold):
    q = ost.bh_correct(gold["bh_p"])
    np.testing.assert_allclose(q, gold["bh_q"], rtol=1e-14, atol=0)
    p = gold["bh_p"].copy()
    p[10] = np.nan
    q2 = ost.bh_correct(p)
    assert np.isnan(q2[10]) and not np.any(np.isnan(np.delete(q2, 10)))


def test_two_group_constructor_rules():
    rng = np.random.default_rng(3)
    n = 400
    x = rng.poisson(2.0, size=(n, 6)).astype(float)
    x[:, 0] = 0
    x[:, 1] = 5
    grouping = rng.integers(0, 2, n)
    x[:, 2] = np.where(grouping == grouping[0], 3.0, 4.0)      # zero variance, different means
    for test in ("t", "rank"):
        res = ost.two_group_test(x, grouping, test)
        assert np.isnan(res["pval"][0]) and res["pval"][1] == 1.0 and res["pval"][2] == 0.0
        res_s = ost.two_group_test(scipy.sparse.csr_matrix(x), grouping, test)
        np.testing.assert_allclose(res_s["pval"][3:], res["pval"][3:], rtol=1e-9)


# ---- NB GLM restatement -----------------------------------------------------------------------------
def _sim(seed=5, n=600, g=12):
    rng = np.random.default_rng(seed)
    cond = rng.integers(0, 2, n).astype(float)
    batch = rng.integers(0, 3, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 3)[None]).astype(float)], axis=1)
    mu = np.exp(rng.uniform(np.log(0.3), np.log(30), g))[None, :] * np.exp(0.4 * cond)[:, None]
    r = rng.uniform(0.7, 3, g)
    y = rng.poisson(rng.gamma(r[None, :], mu / r[None, :])).astype(float)
    return y, xd, np.ones((n, 1))


def test_scale_gradient_matches_finite_differences():
    y, xd, zd = _sim()
    m = onb.NBModel(y, xd, zd, stable=True)
    rng = np.random.default_rng(0)
    a = np.zeros((4, y.shape[1]))
    a[0] = np.log(y.mean(axis=0) + 0.1)
    b = rng.normal(0.3, 0.2, (1, y.shape[1]))
    idx = np.arange(y.shape[1])
    grad, hess = m.scale_grad_hess(a, b, idx)
    h = 1e-5
    f = lambda bb: m.ll_byfeature(a, bb, idx)
    g_fd = (f(b + h) - f(b - h)) / (2 * h)
    h_fd = (m.scale_grad_hess(a, b + h, idx)[0] - m.scale_grad_hess(a, b - h, idx)[0])[:, 0] / (2 * h)
    np.testing.assert_allclose(grad[:, 0], g_fd, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(hess[:, 0, 0], h_fd, rtol=1e-6, atol=1e-6)


def test_stable_and_naive_likelihood_agree():
    y, xd, zd = _sim()
    idx = np.arange(y.shape[1])
    a = np.zeros((4, y.shape[1]))
    a[0] = np.log(y.mean(axis=0) + 0.1)
    b = np.full((1, y.shape[1]), 0.4)
    l1 = onb.NBModel(y, xd, zd, stable=True).ll_byfeature(a, b, idx)
    l2 = onb.NBModel(y, xd, zd, stable=False).ll_byfeature(a, b, idx)
    np.testing.assert_allclose(l1, l2, rtol=1e-12)


def test_brent_and_newton_modes_agree_within_north_star_tolerance():
    y, xd, zd = _sim()
    fa = onb.fit_nb(y, xd, zd, method_b="brent")
    fb = onb.fit_nb(y, xd, zd, method_b="newton")
    assert np.all(fa["error_codes"] & onb.FLAG_CONVERGED) and np.all(fb["error_codes"] & onb.FLAG_CONVERGED)
    np.testing.assert_allclose(fb["a_var"], fa["a_var"], rtol=1e-5, atol=2e-6)
    sd_a = np.sqrt(fa["fisher_inv"][:, 1, 1])
    sd_b = np.sqrt(fb["fisher_inv"][:, 1, 1])
    np.testing.assert_allclose(sd_b, sd_a, rtol=1e-5)
    # the fitted model is a stationary point of the likelihood
    assert np.max(fb["jacobian"]) < 1e-4


def test_committed_nbglm_goldens_are_reproducible(golden_dir):
    for name in ("nbglm_cfg1.npz", "nbglm_batch.npz"):
        gd = np.load(os.path.join(golden_dir, name))
        y = gd["x"].astype(np.float64)
        sub = slice(0, 8)
        fit = onb.fit_nb(y[:, sub], gd["design_loc"], gd["design_scale"], method_b="newton")
        np.testing.assert_allclose(fit["a_var"], gd["newton_a_var"][:, sub], rtol=1e-9, atol=1e-10)
        np.testing.assert_array_equal(fit["niter"], gd["newton_niter"][sub])
        np.testing.assert_array_equal(fit["error_codes"], gd["newton_error_codes"][sub])


def test_rank_deficient_design_raises():
    y, xd, zd = _sim()
    bad = np.concatenate([xd, xd[:, [1]]], axis=1)
    with pytest.raises(ValueError):
        onb.fit_nb(y, bad, zd, method_b="newton")


def test_oracle_fit_reaches_an_independent_maximum_likelihood_estimate():
    """The NB-GLM oracle restates batchglm's schedule, which cannot be run here.  What can be pinned
    independently is WHERE the schedule ends: a generic quasi-Newton maximiser of the NB log-likelihood (scipy, analytic
    gradient written from the model definition only) must arrive at the same coefficients, dispersion and likelihood, and
    the reported Fisher inverse must equal (X' W X)^-1 with W = mu r / (r + mu) evaluated there."""
    import scipy.optimize
    from scipy.special import gammaln, digamma
    rng = np.random.default_rng(42)
    n, g = 600, 8
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 3, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 3)[None])], axis=1).astype(float)
    mu_true = np.exp(rng.uniform(np.log(0.3), np.log(40.0), g))[None] * np.exp(0.4 * cond)[:, None] * np.exp(0.2 * batch)[:, None]
    r_true = rng.uniform(0.7, 5.0, g)
    y = rng.poisson(rng.gamma(r_true[None], mu_true / r_true[None])).astype(np.float64)
    fit = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
    p = xd.shape[1]
    for j in range(g):
        yj = y[:, j]

        def negll(theta):
            a, b = theta[:p], theta[p]
            eta = xd @ a
            mu, r = np.exp(eta), np.exp(b)
            ll = np.sum(gammaln(yj + r) - gammaln(r) - gammaln(yj + 1.0) + r * (b - np.log(r + mu)) + yj * (eta - np.log(r + mu)))
            d_eta = yj - (yj + r) * mu / (r + mu)                               # d ll / d eta_i
            d_b = r * np.sum(digamma(yj + r) - digamma(r) + b - np.log(r + mu) + 1.0 - (yj + r) / (r + mu))
            return -ll, -np.concatenate([xd.T @ d_eta, [d_b]])

        start = np.concatenate([[np.log(yj.mean() + 1e-3)], np.zeros(p - 1), [0.0]])
        opt = scipy.optimize.minimize(negll, start, jac=True, method="BFGS", options={"gtol": 1e-7, "maxiter": 2000})
        ll_opt = -opt.fun
        a_hat, b_hat = fit["a_var"][:, j], fit["b_var"][0, j]
        # same maximum: the schedule stops on a relative likelihood change of 1e-10, BFGS on the gradient norm
        assert abs(fit["log_likelihood"][j] - ll_opt) <= 1e-7 * abs(ll_opt), (j, fit["log_likelihood"][j], ll_opt)
        mu_hat, r_hat = np.exp(xd @ a_hat), np.exp(b_hat)
        w = mu_hat * r_hat / (r_hat + mu_hat)
        cov = np.linalg.inv(xd.T @ (w[:, None] * xd))
        sd = np.sqrt(np.diag(cov))
        assert np.all(np.abs(a_hat - opt.x[:p]) <= 2e-3 * sd + 1e-6), (j, a_hat, opt.x[:p])
        assert abs(b_hat - opt.x[p]) <= 2e-3
        np.testing.assert_allclose(fit["fisher_inv"][j], cov, rtol=1e-8, atol=1e-12)


def test_oracle_fit_with_size_factors_and_scale_model_reaches_the_independent_mle():
    """Same independent pin for the general model: per-cell size factors as offsets and a dispersion model with two
    coefficients ('~ 1 + condition')."""
    import scipy.optimize
    from scipy.special import gammaln, digamma
    rng = np.random.default_rng(7)
    n, g = 500, 5
    cond = rng.integers(0, 2, n)
    xd = np.stack([np.ones(n), cond.astype(float), rng.normal(size=n)], axis=1)
    zd = np.stack([np.ones(n), cond.astype(float)], axis=1)
    sf = rng.uniform(0.5, 2.0, n)
    mu_true = np.exp(rng.uniform(np.log(2.0), np.log(30.0), g))[None] * np.exp(0.5 * cond + 0.2 * xd[:, 2])[:, None] * sf[:, None]
    r_true = np.exp(rng.uniform(0.0, 1.5, g))[None] * np.exp(0.6 * cond)[:, None]
    y = rng.poisson(rng.gamma(r_true, mu_true / r_true)).astype(np.float64)
    fit = onb.fit_nb(y, xd, zd, size_factors=sf, method_b="newton")
    p, ps = xd.shape[1], zd.shape[1]
    for j in range(g):
        yj = y[:, j]

        def negll(theta):
            a, b = theta[:p], theta[p:]
            eta = xd @ a + np.log(sf)
            zeta = zd @ b
            mu, r = np.exp(eta), np.exp(zeta)
            ll = np.sum(gammaln(yj + r) - gammaln(r) - gammaln(yj + 1.0) + r * (zeta - np.log(r + mu)) + yj * (eta - np.log(r + mu)))
            d_eta = yj - (yj + r) * mu / (r + mu)
            d_zeta = r * (digamma(yj + r) - digamma(r) + zeta - np.log(r + mu) + 1.0 - (yj + r) / (r + mu))
            return -ll, -np.concatenate([xd.T @ d_eta, zd.T @ d_zeta])

        start = np.concatenate([[np.log(np.mean(yj / sf) + 1e-3)], np.zeros(p - 1), np.zeros(ps)])
        opt = scipy.optimize.minimize(negll, start, jac=True, method="BFGS", options={"gtol": 1e-7, "maxiter": 3000})
        ll_opt = -opt.fun
        assert abs(fit["log_likelihood"][j] - ll_opt) <= 1e-7 * abs(ll_opt), (j, fit["log_likelihood"][j], ll_opt)
        a_hat, b_hat = fit["a_var"][:, j], fit["b_var"][:, j]
        mu_hat, r_hat = np.exp(xd @ a_hat + np.log(sf)), np.exp(zd @ b_hat)
        w = mu_hat * r_hat / (r_hat + mu_hat)
        cov = np.linalg.inv(xd.T @ (w[:, None] * xd))
        sd = np.sqrt(np.diag(cov))
        assert np.all(np.abs(a_hat - opt.x[:p]) <= 3e-3 * sd + 1e-6), (j, a_hat, opt.x[:p])
        assert np.all(np.abs(b_hat - opt.x[p:]) <= 5e-3), (j, b_hat, opt.x[p:])
        np.testing.assert_allclose(fit["fisher_inv"][j], cov, rtol=1e-8, atol=1e-12)
