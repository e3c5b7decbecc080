"""GPU (-m gpu): the multi-test wrappers (pairwise / versus_rest / partition; reference tests.py:1097-1968,
unit_test/test_pairwise_null.py, test_vsrest.py, test_partition.py re-expressed).  The model-free variants read the
matrix once (K_ingest over all groups + dx_group_pair_tests per comparison); they must agree exactly with the
reference's loop over two-sample tests, which is what the oracle restates."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse
import scipy.stats

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import oracle.stats as ost           # noqa: E402


@pytest.fixture(scope="module")
def de():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffxpy_b200.api as de_
    return de_


def _counts(seed, n=1500, g=40, k=3, noninteger=False):
    rng = np.random.default_rng(seed)
    grouping = rng.integers(0, k, n).astype(str)
    lam = np.exp(rng.uniform(np.log(0.05), np.log(20), g))
    shift = rng.normal(0, 0.4, (k, g))
    code = grouping.astype(int)
    x = rng.poisson(rng.gamma(1.0, lam[None, :] * np.exp(shift[code]))).astype(np.float64)
    x[:, 0] = 0.0               # all-zero gene
    x[:, 1] = 3.0               # constant gene
    if noninteger:
        x[::5, 2] += 0.5        # overflow entries in every group
        x[::9, 3] = -1.0
        x[:, 4] = 80.0 + code   # above the direct bins, differs by group
    return x, grouping


@pytest.mark.parametrize("test", ["rank", "t-test"])
@pytest.mark.parametrize("noninteger", [False, True])
def test_pairwise_model_free_single_pass_matches_two_sample_loop(de, test, noninteger):
    x, grouping = _counts(3, noninteger=noninteger)
    names = ["g%d" % i for i in range(x.shape[1])]
    res = de.test.pairwise(data=x, grouping=grouping, test=test, lazy=False, gene_names=names)
    groups = np.unique(grouping)
    assert res.pval.shape == (3, 3, x.shape[1])
    for i in range(3):
        for j in range(i + 1, 3):
            idx = np.where((grouping == groups[i]) | (grouping == groups[j]))[0]
            ref = ost.two_group_test(x[idx], grouping[idx], "rank" if test == "rank" else "t")
            np.testing.assert_allclose(res.pval[i, j], ref["pval"], rtol=1e-10, atol=1e-300, equal_nan=True)
            np.testing.assert_array_equal(res.pval[j, i], res.pval[i, j])
            np.testing.assert_allclose(res.log_fold_change()[i, j], ref["logfc"], rtol=1e-12, atol=1e-12, equal_nan=True)
            np.testing.assert_array_equal(res.log_fold_change()[j, i], -res.log_fold_change()[i, j])
    # the loop path (keep_full_test_objs forces one two_sample call per pair) gives identical numbers
    loop = de.test.pairwise(data=x, grouping=grouping, test=test, lazy=False, gene_names=names, keep_full_test_objs=True)
    np.testing.assert_array_equal(np.nan_to_num(loop.pval, nan=-1), np.nan_to_num(res.pval, nan=-1))
    np.testing.assert_allclose(loop.log_fold_change(), res.log_fold_change(), rtol=1e-13, atol=0, equal_nan=True)
    summ = res.summary()
    assert list(summ.columns) == ["gene", "pval", "qval", "log2fc", "mean"]
    sp = res.summary_pairs(groups0=[groups[0]], groups1=[groups[1], groups[2]])
    assert sp.shape[0] == x.shape[1]
    # rank p-values against scipy directly for one pair
    if test == "rank" and not noninteger:
        a, b = x[grouping == groups[0]], x[grouping == groups[2]]
        for gi in (5, 17):
            want = scipy.stats.mannwhitneyu(a[:, gi], b[:, gi], use_continuity=True, alternative="two-sided").pvalue
            assert abs(res.pval[0, 2, gi] - want) <= 1e-12 * want


@pytest.mark.parametrize("test", ["rank", "t-test"])
def test_versus_rest_single_pass_matches_loop(de, test):
    x, grouping = _counts(5, k=4, noninteger=True)
    names = ["g%d" % i for i in range(x.shape[1])]
    res = de.test.versus_rest(data=scipy.sparse.csr_matrix(x), grouping=grouping, test=test, gene_names=names)
    loop = de.test.versus_rest(data=x, grouping=grouping, test=test, gene_names=names, keep_full_test_objs=True)
    assert res.pval.shape == (1, 4, x.shape[1])
    np.testing.assert_array_equal(np.nan_to_num(loop.pval, nan=-1), np.nan_to_num(res.pval, nan=-1))
    np.testing.assert_allclose(loop.log_fold_change(), res.log_fold_change(), rtol=1e-13, atol=0, equal_nan=True)
    groups = np.unique(grouping)
    for i, g1 in enumerate(groups):
        ref = ost.two_group_test(x, np.where(grouping == g1, "group", "rest"), "rank" if test == "rank" else "t")
        np.testing.assert_allclose(res.pval_group(g1), ref["pval"], rtol=1e-10, atol=1e-300, equal_nan=True)
        np.testing.assert_allclose(res.log_fold_change_group(g1), ref["logfc"], rtol=1e-12, atol=1e-12, equal_nan=True)
    assert res.summary_group(groups[1]).shape[0] == x.shape[1]
    assert np.all(res.qval[np.isfinite(res.qval)] <= 1.0)


def test_pairwise_ztest_one_fit(de):
    """tests.py:1224-1262 + det_pair.py:373-387: one fit '~ 0 + group', p = two_coef_z_test of the group coefficients."""
    rng = np.random.default_rng(7)
    n, g, k = 3000, 60, 3
    grouping = rng.integers(0, k, n).astype(str)
    mu = np.exp(rng.uniform(np.log(5), np.log(50), g))
    x = rng.poisson(rng.gamma(2.0, mu[None, :] / 2.0, size=(n, g))).astype(np.float64)      # null: no group effect
    res = de.test.pairwise(data=x, grouping=grouping, test="z-test", noise_model="nb", gene_names=["g%d" % i for i in range(g)])
    est = res.model_estim
    assert est.a_var.shape == (k, g)
    sd = np.sqrt(np.diagonal(est.fisher_inv, axis1=-2, axis2=-1).T)
    for i in range(k):
        for j in range(i + 1, k):
            want = ost.two_coef_z_test(est.a_var[i], est.a_var[j], sd[i], sd[j])
            np.testing.assert_allclose(res.pval[i, j], want, rtol=1e-12, atol=1e-300)
            np.testing.assert_array_equal(res.pval[j, i], res.pval[i, j])
            np.testing.assert_array_equal(res.log_fold_change()[i, j], est.a_var[j] - est.a_var[i])
    iu = np.triu_indices(k, 1)
    pv = res.pval[iu].reshape(-1)
    assert scipy.stats.kstest(pv, "uniform").pvalue > 0.01, "pairwise z-test p-values are not calibrated under the null"
    assert res.summary().shape[0] == g
    assert res.qval.shape == res.pval.shape
    by_test = de.test.pairwise(data=x, grouping=grouping, test="z-test", noise_model="nb", pval_correction="by_test",
                               gene_names=["g%d" % i for i in range(g)])
    assert by_test.qval.shape == res.pval.shape


def test_partition_equals_tests_on_subsets(de):
    rng = np.random.default_rng(9)
    n, g = 2400, 30
    part = rng.integers(0, 2, n).astype(str)
    cond = rng.integers(0, 2, n).astype(str)
    mu = np.exp(rng.uniform(np.log(2), np.log(40), g))
    x = rng.poisson(rng.gamma(1.5, mu[None, :] / 1.5, size=(n, g))).astype(np.float64)
    sd = pd.DataFrame({"part": part, "condition": cond})
    names = ["g%d" % i for i in range(g)]
    p = de.test.partition(data=x, parts="part", gene_names=names, sample_description=sd)
    res = p.wald(factor_loc_totest="condition", formula_loc="~ 1 + condition")
    assert res.pval.shape == (1, 2, g)
    for i, lab in enumerate(np.unique(part)):
        idx = np.where(part == lab)[0]
        single = de.test.wald(data=x[idx], formula_loc="~ 1 + condition", factor_loc_totest="condition",
                              sample_description=sd.iloc[idx], gene_names=names)
        np.testing.assert_array_equal(res.pval[0, i], single.pval)
        np.testing.assert_array_equal(res.test_of(lab).pval, single.pval)
    r2 = p.rank_test(grouping="condition")
    for i, lab in enumerate(np.unique(part)):
        idx = np.where(part == lab)[0]
        ref = ost.two_group_test(x[idx], cond[idx], "rank")
        np.testing.assert_allclose(r2.pval[0, i], ref["pval"], rtol=1e-10, equal_nan=True)
    assert list(res.summary().columns) == ["gene", "pval", "qval", "log2fc", "mean"]
    t2 = p.two_sample(grouping="condition", test="t-test")
    assert t2.pval.shape == (1, 2, g)


# ---------------------------------------------------------------------------------------------------------
# continuous_1d (reference unit_test/test_continuous_null.py, test_continuous_de.py re-expressed)
# ---------------------------------------------------------------------------------------------------------
def _continuous_data(seed, n=2000, g=100, de_frac=0.5):
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 1, n)
    base = rng.uniform(np.log(5), np.log(50), g)
    amp = np.where(np.arange(g) < int(de_frac * g), rng.uniform(0.6, 1.2, g), 0.0)
    eta = base[None, :] + amp[None, :] * np.sin(2 * np.pi * t)[:, None]
    x = rng.poisson(rng.gamma(2.0, np.exp(eta) / 2.0)).astype(np.float64)
    sd = pd.DataFrame({"pseudotime": t, "batch": rng.integers(0, 2, n).astype(str)})
    return x, sd, amp > 0


def test_continuous_1d_null_calibration_and_power(de):
    x, sd, is_de = _continuous_data(11, de_frac=0.0)
    names = ["g%d" % i for i in range(x.shape[1])]
    res = de.test.continuous_1d(data=x, continuous="pseudotime", df=5, formula_loc="~ 1 + pseudotime",
                                formula_scale="~ 1", factor_loc_totest="pseudotime", test="wald",
                                sample_description=sd, gene_names=names)
    assert scipy.stats.kstest(res.pval, "uniform").pvalue > 0.01, "continuous_1d Wald p-values are not calibrated"
    assert res._de_test.model_estim.a_var.shape[0] == 6                      # intercept + df spline coefficients
    x, sd, is_de = _continuous_data(12, de_frac=0.5)
    res = de.test.continuous_1d(data=x, continuous="pseudotime", df=5, formula_loc="~ 1 + pseudotime + batch",
                                formula_scale="~ 1", factor_loc_totest="pseudotime", test="wald",
                                sample_description=sd, gene_names=names)
    q = res.qval
    assert np.mean(q[is_de] < 0.05) >= 0.5 and np.mean(q[~is_de] < 0.05) <= 0.1
    summ = res.summary()
    assert list(summ.columns[:6]) == ["gene", "pval", "qval", "log2fc", "mean", "zero_mean"]
    # fold change of the fitted curve: GPU scan == per-gene numpy evaluation of the reference recipe
    genes = [0, 3, 77]
    vals = res._continuous_model(idx=np.array(genes))
    want = (np.log(vals.max(axis=0)) - np.log(vals.min(axis=0))) / np.log(2)
    np.testing.assert_allclose(res.log2_fold_change(genes=genes), want, rtol=1e-10)
    np.testing.assert_allclose(summ["log2fc"].values[genes], want, rtol=1e-10)
    np.testing.assert_array_equal(res.argmax(genes), sd["pseudotime"].values[vals.argmax(axis=0)])
    np.testing.assert_allclose(res.max(genes), vals.max(axis=0), rtol=1e-12)
    assert np.mean(summ["log2fc"].values[is_de]) > 3 * np.mean(summ["log2fc"].values[~is_de])
    t_eval, mu = res._continuous_interpolation(idx=0)
    assert t_eval.shape == (100,) and mu.shape == (100, 1)


def test_continuous_1d_lrt_and_errors(de):
    x, sd, is_de = _continuous_data(13, n=1500, g=40, de_frac=0.5)
    names = ["g%d" % i for i in range(x.shape[1])]
    res = de.test.continuous_1d(data=x, continuous="pseudotime", df=4, formula_loc="~ 1 + pseudotime + batch",
                                formula_scale="~ 1", factor_loc_totest="pseudotime", test="lrt",
                                sample_description=sd, gene_names=names)
    assert res._de_test.reduced_estim.a_var.shape[0] == 2 and res._de_test.full_estim.a_var.shape[0] == 6
    assert np.mean(res.qval[is_de] < 0.05) >= 0.5 and np.mean(res.qval[~is_de] < 0.05) <= 0.15
    with pytest.raises(ValueError):
        de.test.continuous_1d(data=x, continuous="nope", df=4, formula_loc="~ 1 + nope", factor_loc_totest="nope",
                              sample_description=sd, gene_names=names)
    with pytest.raises(ValueError):
        de.test.continuous_1d(data=x, continuous="batch", df=4, formula_loc="~ 1 + batch", factor_loc_totest="batch",
                              sample_description=sd, gene_names=names)
    with pytest.raises(ValueError):
        de.test.continuous_1d(data=x, continuous="pseudotime", df=4, spline_basis="xx", formula_loc="~ 1 + pseudotime",
                              factor_loc_totest="pseudotime", sample_description=sd, gene_names=names)


@pytest.mark.parametrize("basis", ["cr", "cc"])
def test_continuous_1d_cubic_regression_splines(de, basis):
    """Natural / cyclic cubic regression spline bases (tests.py:2201-2210): df centred columns, calibrated under the null,
    powered on genes that change along the covariate."""
    x, sd, is_de = _continuous_data(21, n=1500, g=60, de_frac=0.0)
    names = ["g%d" % i for i in range(x.shape[1])]
    res = de.test.continuous_1d(data=x, continuous="pseudotime", df=4, spline_basis=basis, formula_loc="~ 1 + pseudotime",
                                formula_scale="~ 1", factor_loc_totest="pseudotime", test="wald",
                                sample_description=sd, gene_names=names)
    assert res._de_test.model_estim.a_var.shape[0] == 5                      # intercept + df spline coefficients
    assert scipy.stats.kstest(res.pval, "uniform").pvalue > 0.01
    x, sd, is_de = _continuous_data(22, n=1500, g=60, de_frac=0.5)
    res = de.test.continuous_1d(data=x, continuous="pseudotime", df=4, spline_basis=basis, formula_loc="~ 1 + pseudotime + batch",
                                formula_scale="~ 1", factor_loc_totest="pseudotime", test="lrt",
                                sample_description=sd, gene_names=names)
    assert np.mean(res.qval[is_de] < 0.05) >= 0.5          # one full sine period: inside both spaces
    assert np.mean(res.qval[~is_de] < 0.05) <= 0.15


def test_gene_sharded_estimator_over_nccl_matches_single_process():
    """SURVEY 8e on real GPUs (needs >= 2): tools/dist_check.py under torchrun -- every rank fits its gene range, one
    NCCL all-gather of the per-gene rows, result bit-identical to the single-process run."""
    import os
    import subprocess
    import sys
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(root, "tools", "dist_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dist_check OK" in out.stdout and "peer exchange OK" in out.stdout, \
        out.stdout[-2000:] + out.stderr[-2000:]
