"""CPU: host-side mirror of the reference interface -- design / constraint builders, argument checks, sharding."""
import numpy as np
import pandas as pd
import pytest

import diffxpy_b200.api as de
from diffxpy_b200 import dist as dxdist
from diffxpy_b200.data import design
from diffxpy_b200.estimator import InputDataGLM, group_rows
from diffxpy_b200.testing import utils as tutils


def _sd(n=24):
    return pd.DataFrame({
        "condition": np.array(["0", "1"] * (n // 2)),
        "batch": np.repeat(np.array(["a", "b", "c", "d"]), n // 4),
        "numeric1": np.linspace(0, 1, n),
    })


def test_design_matrix_names_and_slices():
    dm = design.design_matrix(_sd(), "~ 1 + condition + batch")
    assert dm.design_info.column_names == ["Intercept", "condition[T.1]", "batch[T.b]", "batch[T.c]", "batch[T.d]"]
    assert dm.design_info.slice("condition") == slice(1, 2)
    assert dm.design_info.slice("batch") == slice(2, 5)
    assert np.linalg.matrix_rank(np.asarray(dm)) == 5
    dm0 = design.design_matrix(_sd(), "~ 0 + batch")
    assert dm0.design_info.column_names == ["batch[a]", "batch[b]", "batch[c]", "batch[d]"]
    np.testing.assert_array_equal(np.asarray(dm0).sum(axis=1), 1)


def test_numeric_covariates_get_one_coefficient_each():
    # unit_test/test_numeric_covar.py:31-41
    sd = _sd()
    sd["numeric2"] = sd["numeric1"] ** 2
    names = tutils.preview_coef_names(sd, "~ 1 + condition + numeric1 + numeric2", as_numeric=["numeric1", "numeric2"])
    assert len(names) == 4 and names[-2:] == ["numeric1", "numeric2"]


def test_constraints_from_dict_and_string():
    # unit_test/test_constrained.py: batches nested in conditions, sum-to-zero per condition
    sd = pd.DataFrame({"condition": np.repeat(["c0", "c1"], 8), "batch": np.repeat(["b0", "b1", "b2", "b3"], 4)})
    dmat, names, cmat, terms = design.constraint_system_from_star(sample_description=sd, formula="~ 1 + condition + batch",
                                                                 constraints={"batch": "condition"})
    assert names == ["Intercept", "condition[T.c1]", "batch[b0]", "batch[b1]", "batch[b2]", "batch[b3]"]
    assert cmat.shape == (6, 4)
    assert np.linalg.matrix_rank(np.asarray(dmat)) < 6 and np.linalg.matrix_rank(np.asarray(dmat) @ cmat) == 4
    # first coefficient of every constraint is the dependent one: b0 = -b1, b2 = -b3
    np.testing.assert_array_equal(cmat[2], [0, 0, -1, 0])
    np.testing.assert_array_equal(cmat[4], [0, 0, 0, -1])
    cm2 = design.constraint_matrix_from_string(dmat, names, ["batch[b0]+batch[b1]=0", "batch[b2]+batch[b3]=0"])
    np.testing.assert_array_equal(cm2, cmat)


def test_rank_deficient_design_raises_value_error():
    # unit_test/test_single_fullrank.py:40-52
    sd = pd.DataFrame({"condition": np.repeat(["c0", "c1"], 8), "batch": np.repeat(["b0", "b1", "b2", "b3"], 4)})
    dm = design.design_matrix(sd, "~ 1 + condition + batch")
    with pytest.raises(ValueError):
        InputDataGLM(np.zeros((16, 3)), dm, np.ones((16, 1)))
    d2, names, cmat, _ = design.constraint_system_from_star(sample_description=sd, formula="~ 1 + condition + batch",
                                                            constraints={"batch": "condition"})
    InputDataGLM(np.zeros((16, 3)), d2, np.ones((16, 1)), constraints_loc=cmat)   # must not raise


def test_wald_argument_errors_match_reference():
    x = np.zeros((24, 3))
    sd = _sd()
    with pytest.raises(ValueError):
        de.test.wald(x, formula_loc=None, sample_description=sd, gene_names=list("abc"))
    with pytest.raises(ValueError):
        de.test.wald(x, formula_loc="~1+condition", sample_description=sd, gene_names=list("abc"))   # nothing to test
    with pytest.raises(ValueError):
        de.test.wald(x, formula_loc="~1+condition", coef_to_test="nope", sample_description=sd, gene_names=list("abc"))
    with pytest.raises(ValueError):
        de.test.wald(x, formula_loc="~1+condition", factor_loc_totest="condition", sample_description=sd,
                     gene_names=list("abc"), backend="tf1")
    with pytest.raises(ValueError):
        de.test.wald(x, formula_loc="~1+condition", factor_loc_totest="condition", sample_description=sd)  # gene names


def test_group_rows_and_split_x():
    m = np.array([[1, 0.0], [1, 1], [1, -0.0], [1, 1]])
    u, inv = group_rows(m)
    assert u.shape[0] == 2 and inv[0] == inv[2] and inv[1] == inv[3]
    x = np.arange(12).reshape(6, 2)
    x0, x1 = tutils.split_x(x, np.array(["b", "a", "b", "a", "a", "b"]))
    np.testing.assert_array_equal(x0, x[[0, 2, 5]])      # first appearance order


def test_gene_ranges_cover_and_balance():
    r = dxdist.gene_ranges(10, 3)
    assert r[0][0] == 0 and r[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    w = np.array([100, 1, 1, 1, 1, 1, 1, 100, 1, 1])
    r = dxdist.gene_ranges(10, 2, w)
    loads = [w[a:b].sum() for a, b in r]
    assert r[0][0] == 0 and r[-1][1] == 10 and abs(loads[0] - loads[1]) <= 100
    assert dxdist.gene_ranges(5, 1) == [(0, 5)]


def test_bs_basis_matches_splev_recipe_and_formula_rewrite():
    """data/splines.py restates patsy's bs(x, df, degree=3, include_intercept=False) (tests.py:2196-2200)."""
    from scipy.interpolate import splev
    from diffxpy_b200.data.splines import bs, bs_knots
    from diffxpy_b200.testing.tests import _replace_continuous
    rng = np.random.default_rng(0)
    x = rng.uniform(-2, 3, 400)
    for df in (3, 5, 8):
        b = bs(x, df)
        assert b.shape == (400, df)
        k = bs_knots(x, df)
        assert len(k) == df + 1 + 4 and k[0] == x.min() and k[-1] == x.max()
        n = len(k) - 4
        ref = np.stack([splev(x, (k, np.eye(n)[i], 3)) for i in range(n)], axis=1)[:, 1:]
        np.testing.assert_allclose(b, ref, atol=1e-14)
        np.testing.assert_allclose(np.quantile(x, np.linspace(0, 1, df - 3 + 2)[1:-1]), k[4:-4])
    assert _replace_continuous("~ 1 + pt + batch + pt:cond", "pt", ["pt0", "pt1"]) == \
        "~ 1 + pt0 + pt1 + batch + pt0:cond + pt1:cond"
    assert _replace_continuous("~ 1", "pt", ["pt0"]) == "~ 1"


@pytest.mark.parametrize("cyclic", [False, True])
def test_cubic_regression_spline_bases_have_the_defining_properties(cyclic):
    """cr / cc of continuous_1d (reference: patsy cr / cc with constraints='center', tests.py:2201-2210).  patsy cannot be
    imported here, so the restatement is pinned by what defines the construction: knots at quantiles of the unique sample
    values, a cardinal basis of C2 piecewise cubics with natural (cr) or periodic (cc) boundary behaviour, and the
    absorbed centering constraint.  These properties fix the spanned space, hence the fit and every test statistic of a
    model with an intercept."""
    from diffxpy_b200.data.splines import crs, crs_knots, crs_free_basis
    rng = np.random.default_rng(3)
    x = np.round(rng.uniform(-2.0, 7.0, 400), 2)              # repeated values: knots come from the unique ones
    df = 5
    knots = crs_knots(x, df, cyclic=cyclic)
    ux = np.unique(x)
    n_inner = df - (1 if cyclic else 2) + 1
    assert knots.size == n_inner + 2 and knots[0] == ux[0] and knots[-1] == ux[-1]
    np.testing.assert_allclose(knots[1:-1], np.percentile(ux, np.linspace(0, 100, n_inner + 2)[1:-1]))
    # cardinal: value 1 at the own knot, 0 at the others
    at_knots = crs_free_basis(knots[:-1] if cyclic else knots, knots, cyclic)
    np.testing.assert_allclose(at_knots, np.eye(at_knots.shape[0]), atol=1e-12)
    # C2 piecewise cubic: third differences of each piece vanish, value / first / second derivative continuous at the knots
    eps = 1e-4
    for t in knots[1:-1]:
        left = crs_free_basis(np.array([t - 2 * eps, t - eps, t]), knots, cyclic)
        right = crs_free_basis(np.array([t, t + eps, t + 2 * eps]), knots, cyclic)
        d1l, d1r = (left[2] - left[1]) / eps, (right[1] - right[0]) / eps
        d2l, d2r = (left[2] - 2 * left[1] + left[0]) / eps ** 2, (right[2] - 2 * right[1] + right[0]) / eps ** 2
        np.testing.assert_allclose(d1l, d1r, atol=5e-3)
        np.testing.assert_allclose(d2l, d2r, atol=5e-2)
    mid = 0.5 * (knots[2] + knots[3])
    pts = crs_free_basis(mid + eps * np.arange(5), knots, cyclic)
    np.testing.assert_allclose(np.diff(pts, n=4, axis=0), 0.0, atol=1e-9)      # cubic inside an interval
    if cyclic:
        # periodic: value, first and second derivative agree across the seam; points outside are wrapped
        span = knots[-1] - knots[0]
        a = crs_free_basis(np.array([knots[0], knots[0] + eps, knots[0] + 2 * eps]), knots, True)
        b = crs_free_basis(np.array([knots[-1] - 2 * eps, knots[-1] - eps, knots[-1] - 1e-12]), knots, True)
        np.testing.assert_allclose(a[0], b[2], atol=1e-6)
        np.testing.assert_allclose((a[1] - a[0]) / eps, (b[2] - b[1]) / eps, atol=5e-3)
        np.testing.assert_allclose((a[2] - 2 * a[1] + a[0]) / eps ** 2, (b[2] - 2 * b[1] + b[0]) / eps ** 2, atol=5e-2)
        np.testing.assert_allclose(crs_free_basis(np.array([1.3 + span]), knots, True), crs_free_basis(np.array([1.3]), knots, True), atol=1e-12)
    else:
        # natural: second derivative zero at the boundary knots, linear continuation beyond them
        for t, sgn in ((knots[0], 1.0), (knots[-1], -1.0)):
            q = crs_free_basis(np.array([t, t + sgn * eps, t + 2 * sgn * eps]), knots, False)
            np.testing.assert_allclose((q[2] - 2 * q[1] + q[0]) / eps ** 2, 0.0, atol=5e-2)
        out = crs_free_basis(knots[-1] + np.array([0.5, 1.0, 1.5]), knots, False)
        np.testing.assert_allclose(np.diff(out, n=2, axis=0), 0.0, atol=1e-10)
    # centred basis: df columns, zero column means over the sample, full rank together with the intercept,
    # and the same transform re-applied to new points spans the same functions
    basis, k2, constraint = crs(x, df, cyclic=cyclic)
    assert basis.shape == (x.size, df)
    np.testing.assert_allclose(basis.mean(axis=0), 0.0, atol=1e-12)
    full = np.hstack([np.ones((x.size, 1)), basis])
    assert np.linalg.matrix_rank(full) == df + 1
    free = crs_free_basis(x, knots, cyclic)
    resid = free - full @ np.linalg.lstsq(full, free, rcond=None)[0]
    np.testing.assert_allclose(resid, 0.0, atol=1e-9)         # span{1, centred basis} == span{free basis}
    again, _, _ = crs(x[:50], df, cyclic=cyclic, knots=k2, constraint=constraint)
    np.testing.assert_allclose(again, basis[:50], atol=1e-12)


def test_formula_coding_follows_patsy_redundancy_rules():
    """Interaction terms, 'C(x)' and term order as patsy builds them (the reference hands its formulas to patsy through
    batchglm's design_matrix, utils.py:127-176).  The expected column names are patsy's documented results for these
    formulas ("Redundancy and categorical factors" in its formula documentation); every design must have full rank."""
    rng = np.random.default_rng(0)
    n = 80
    sd = pd.DataFrame({"a": rng.choice(["a1", "a2"], n), "b": rng.choice(["b1", "b2"], n), "c": rng.integers(0, 3, n),
                       "x": rng.normal(size=n), "z": rng.normal(size=n)})
    cat = [True, True, False, False, False]
    want = {
        "~ 1 + a + b": ["Intercept", "a[T.a2]", "b[T.b2]"],
        "~ 0 + a + b": ["a[a1]", "a[a2]", "b[T.b2]"],
        "~ 1 + a:b": ["Intercept", "b[T.b2]", "a[T.a2]:b[b1]", "a[T.a2]:b[b2]"],
        "~ 0 + a:b": ["a[a1]:b[b1]", "a[a2]:b[b1]", "a[a1]:b[b2]", "a[a2]:b[b2]"],
        "~ 1 + a + a:b": ["Intercept", "a[T.a2]", "a[a1]:b[T.b2]", "a[a2]:b[T.b2]"],
        "~ 1 + a*b": ["Intercept", "a[T.a2]", "b[T.b2]", "a[T.a2]:b[T.b2]"],
        "~ 1 + a:b + a": ["Intercept", "a[T.a2]", "a[a1]:b[T.b2]", "a[a2]:b[T.b2]"],          # main effects before interactions
        "~ 1 + a:x": ["Intercept", "a[a1]:x", "a[a2]:x"],
        "~ 1 + x + a:x": ["Intercept", "x", "a[T.a2]:x"],
        "~ 1 + x + a": ["Intercept", "a[T.a2]", "x"],                                          # categorical terms first
        "~ 1 + x + a + a:x + z": ["Intercept", "a[T.a2]", "x", "a[T.a2]:x", "z"],
        "~ 1 + C(c)": ["Intercept", "C(c)[T.1]", "C(c)[T.2]"],
        "~ 1 + c": ["Intercept", "c"],
    }
    for formula, names in want.items():
        dm = design.design_matrix(sd, formula, as_categorical=cat)
        assert list(dm.design_info.column_names) == names, formula
        arr = np.asarray(dm)
        assert np.linalg.matrix_rank(arr) == arr.shape[1], formula
        sl = dm.design_info.term_name_slices
        assert sum(s.stop - s.start for s in sl.values()) == arr.shape[1]
    # the interaction columns are products of the factor codings
    dm = design.design_matrix(sd, "~ 1 + a + a:b", as_categorical=cat)
    arr = np.asarray(dm)
    np.testing.assert_array_equal(arr[:, 2], ((sd["a"] == "a1") & (sd["b"] == "b2")).values.astype(float))
    np.testing.assert_array_equal(arr[:, 3], ((sd["a"] == "a2") & (sd["b"] == "b2")).values.astype(float))
    dm = design.design_matrix(sd, "~ 1 + a:x", as_categorical=cat)
    np.testing.assert_allclose(np.asarray(dm)[:, 1], (sd["a"] == "a1").values * sd["x"].values)


def test_lrt_locations_and_scales_tables():
    """det.py:606-648 on stub estimators (no GPU): fitted mean / dispersion for every distinct factor combination."""
    from types import SimpleNamespace
    from diffxpy_b200.testing.det import DifferentialExpressionTestLRT
    rng = np.random.default_rng(1)
    sd = pd.DataFrame({"condition": rng.choice(["c0", "c1"], 40), "batch": rng.choice(["b0", "b1", "b2"], 40)})
    dm = design.design_matrix(sd, "~ 1 + condition + batch")
    ds = design.design_matrix(sd, "~ 1 + condition")
    a = rng.normal(size=(dm.shape[1], 3))
    b = rng.normal(size=(ds.shape[1], 3))
    model = SimpleNamespace(a=a, b=b, inverse_link_loc=np.exp, inverse_link_scale=np.exp)
    est = SimpleNamespace(model=model, input_data=SimpleNamespace(design_loc=np.asarray(dm), design_scale=np.asarray(ds),
                                                                   features=["g0", "g1", "g2"]))
    red = design.design_matrix(sd, "~ 1 + batch")
    test = DifferentialExpressionTestLRT(sample_description=sd, full_design_loc_info=dm.design_info, full_estim=est,
                                         reduced_design_loc_info=red.design_info, reduced_estim=est)
    loc = test.locations()
    assert loc.shape == (6, 3) and list(loc.index.names) == ["condition", "batch"] and list(loc.columns) == ["g0", "g1", "g2"]
    row = np.asarray(dm)[(sd["condition"] == "c1").values & (sd["batch"] == "b2").values][0]
    np.testing.assert_allclose(loc.loc[("c1", "b2")].values, np.exp(row @ a))
    sc = test.scales()
    assert sc.shape == (2, 3)
    np.testing.assert_allclose(np.sort(sc["g1"].values), np.sort(np.exp(np.array([[1, 0], [1, 1]]) @ b)[:, 1]))


def test_hypergeom_test_matches_scipy_survival_function():
    import scipy.stats
    from diffxpy_b200.stats.stats import hypergeom_test      # stats.py:329-355
    inter, refs = np.array([5, 1, 0, 12]), np.array([30, 40, 10, 25])
    got = hypergeom_test(inter, 20, refs, 1000)
    want = scipy.stats.hypergeom(M=1000, n=refs, N=20).sf(inter - 1)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-15)
    assert got[2] == 1.0


def test_existing_categorical_columns_keep_their_level_order():
    sd = pd.DataFrame({"condition": pd.Categorical(["treated", "ctrl", "treated", "ctrl", "other"],
                                                   categories=["treated", "other", "ctrl", "unused"])})
    dm = design.design_matrix(sd, "~ 1 + condition")
    assert list(dm.design_info.column_names) == ["Intercept", "condition[T.other]", "condition[T.ctrl]"]
    np.testing.assert_array_equal(np.asarray(dm)[:, 2], [0, 1, 0, 1, 0])
    sd2 = pd.DataFrame({"condition": ["treated", "ctrl", "treated", "ctrl", "other"]})
    assert list(design.design_matrix(sd2, "~ 1 + condition").design_info.column_names) == \
        ["Intercept", "condition[T.other]", "condition[T.treated]"]


def test_anndata_like_inputs_supply_gene_names_and_sample_description():
    """unit_test/test_data_types.py:15-124 feed AnnData / AnnData.raw; anndata is not installed here, so the host logic
    accepts anything that looks like it (X, obs, var_names)."""
    from types import SimpleNamespace
    from diffxpy_b200.testing import utils
    x = np.arange(12.0).reshape(4, 3)
    adata = SimpleNamespace(X=x, obs=pd.DataFrame({"condition": ["a", "b", "a", "b"]}), var_names=pd.Index(["g1", "g2", "g3"]))
    assert list(utils.parse_gene_names(adata, None)) == ["g1", "g2", "g3"]
    assert list(utils.parse_gene_names(x, ["u", "v", "w"])) == ["u", "v", "w"]
    with pytest.raises(ValueError):
        utils.parse_gene_names(x, None)
    sd = utils.parse_sample_description(adata, None)
    assert list(sd["condition"]) == ["a", "b", "a", "b"]
    with pytest.raises(ValueError):
        utils.parse_sample_description(x, None)
    with pytest.raises(AssertionError):
        utils.parse_sample_description(x, pd.DataFrame({"condition": ["a", "b"]}))
    np.testing.assert_array_equal(utils.parse_grouping(adata, None, "condition"), ["a", "b", "a", "b"])
    sf = utils.parse_size_factors(np.array([1.0, 2.0, 0.5, 1.5]), adata, sd)
    assert sf.shape[0] == 4
    with pytest.raises(AssertionError):
        utils.parse_size_factors(np.array([1.0, 0.0, 0.5, 1.5]), adata, sd)


def test_group_rows_many_small_integer_columns_match_the_exact_grouping():
    """0 / 1 / 2-valued designs differ only in the top bits of their float64 patterns; the row hash must still separate
    them (a weak hash silently falls back to the slow exact path -- or worse, merges rows)."""
    from diffxpy_b200.estimator import group_rows, _group_rows_exact
    rng = np.random.default_rng(5)
    n = 20000
    cond, batch, extra = rng.integers(0, 2, n), rng.integers(0, 8, n), rng.integers(0, 3, n)
    onehot = (batch[:, None] == np.arange(1, 8)[None, :]).astype(float)
    mat = np.concatenate([np.ones((n, 1)), cond[:, None], onehot, onehot, extra[:, None]], axis=1).astype(float)
    mat[::7, 1] *= -0.0          # -0.0 and 0.0 are the same design value
    uniq, inv = group_rows(mat)
    uniq_e, inv_e = _group_rows_exact(np.ascontiguousarray(mat + 0.0))
    assert uniq.shape[0] == 48
    np.testing.assert_array_equal(uniq, uniq_e)
    np.testing.assert_array_equal(inv, inv_e)
    np.testing.assert_array_equal(uniq[inv], mat + 0.0)


def test_bench_workload_helpers():
    """bench.py --workload lrt / spline / rank: the host-side pieces that run without a GPU."""
    import importlib
    bw = importlib.import_module("bench_workloads")
    assert set(bw.METRICS) == set(bw.DEFAULT_STEPS) == {"lrt", "spline", "rank"}
    xd, sf = bw.spline_design(5000)
    assert xd.shape == (5000, 12) and np.linalg.matrix_rank(xd) == 12          # configs[3]: 11 spline columns + intercept
    assert np.all(xd[:, 0] == 1.0) and np.all(xd[:, 1:] >= 0.0) and np.all((sf >= 0.5) & (sf <= 1.5))
    peak, src = bw._peaks()
    assert peak > 1000.0 and "GB/s" in src or "MEASURED_PEAKS" in src


def test_result_object_clean_logs_and_threshold_rules():
    """det.py:120-205 behaviour of the result base class: cleaned log10 p-values and the summary() row filters."""
    from diffxpy_b200.testing.det import _DifferentialExpressionTest as Base
    p = np.array([0.0, 1e-320, 1e-40, 1e-5, 0.5, 1.0, np.nan, 2.0])
    np.testing.assert_allclose(Base._log10_floor(p, -30), [-30, -30, -30, -5, np.log10(0.5), 0, 0, 0], rtol=1e-15)
    np.testing.assert_allclose(Base._log10_floor(p, -50)[:3], [-50, -50, -40], rtol=1e-15)

    class T(Base):
        gene_ids = None
        x = None

        def log_fold_change(self, base=np.e, **kwargs):
            return None

        def summary(self, **kwargs):
            return None
    df = pd.DataFrame({"gene": list("abcdef"), "qval": [0.01, np.nan, 0.2, 0.04, 0.9, 0.03],
                       "log2fc": [2.0, 3.0, -2.0, 0.1, -3.0, 1.0], "mean": [1.0, 5.0, 0.1, 2.0, 3.0, 0.5]})
    t = T()
    assert list(t._threshold_summary(df, qval_thres=0.05)["gene"]) == ["a", "d", "f"]          # NaN q-values drop out
    assert list(t._threshold_summary(df, fc_upper_thres=2.0)["gene"]) == ["a", "b", "f"]
    assert list(t._threshold_summary(df, fc_lower_thres=0.5)["gene"]) == ["c", "e"]
    assert list(t._threshold_summary(df, fc_upper_thres=4.0, fc_lower_thres=0.25)["gene"]) == ["a", "b", "c", "e"]
    assert list(t._threshold_summary(df, qval_thres=0.05, fc_upper_thres=2.0, mean_thres=0.8)["gene"]) == ["a"]
    with pytest.raises(AssertionError):
        t._threshold_summary(df, fc_lower_thres=-1.0)


def test_float64_values_rounded_to_float32_are_reported(caplog):
    """device.py stores float32 values: non-representable float64 input (log-normalised data) is reported once."""
    import logging
    from diffxpy_b200 import device
    device._WARNED_F32[0] = False
    with caplog.at_level(logging.WARNING, logger="diffxpy"):
        device._warn_if_rounded_to_float32(np.arange(1000, dtype=np.float64))          # counts: exact
        assert not caplog.records
        device._warn_if_rounded_to_float32(np.log1p(np.arange(1000, dtype=np.float64) / 3.0))
        assert len(caplog.records) == 1 and "float32" in caplog.records[0].getMessage()
        device._warn_if_rounded_to_float32(np.log1p(np.arange(1000, dtype=np.float64) / 7.0))
        assert len(caplog.records) == 1          # once per process
    device._WARNED_F32[0] = False


def test_join_groupings_equals_group_rows_of_the_concatenation():
    """Estimator.initialize derives the joint design groups from the groupings of the location and the scale design."""
    from diffxpy_b200.estimator import group_rows, join_groupings
    rng = np.random.default_rng(9)
    n = 30000
    cond, batch, extra = rng.integers(0, 2, n), rng.integers(0, 8, n), rng.integers(0, 3, n)
    onehot = (batch[:, None] == np.arange(1, 8)[None, :]).astype(float)
    loc = np.concatenate([np.ones((n, 1)), cond[:, None], onehot], axis=1).astype(float)
    scale = np.concatenate([np.ones((n, 1)), onehot[:, :3], -extra[:, None]], axis=1).astype(float)          # negative values too
    third = rng.choice([0.25, -1.5, 3.0], n)[:, None]
    for mats in ([loc, scale], [scale, loc], [loc, scale, third], [loc[:, :1], scale[:, :1]]):
        want_u, want_inv = group_rows(np.concatenate(mats, axis=1))
        got_u, got_inv = join_groupings([group_rows(m) for m in mats])
        np.testing.assert_array_equal(got_u, want_u)
        np.testing.assert_array_equal(got_inv, want_inv)
    big = [(np.arange(5000.0)[:, None], rng.integers(0, 5000, n)), (np.arange(5000.0)[:, None], rng.integers(0, 5000, n))]
    assert join_groupings(big) is None          # too many index tuples for a counting pass: the caller hashes the rows


def test_factor_coding_one_pass_matches_the_definition():
    """design_matrix codes a factor with one hash pass: same columns, names and reference level as comparing the column
    with every sorted level in turn (patsy treatment coding); missing values give all-zero rows."""
    from diffxpy_b200.data import design as dd
    rng = np.random.default_rng(11)
    n = 5000
    sd = pd.DataFrame({"condition": rng.integers(0, 2, n).astype(str), "batch": rng.integers(0, 8, n).astype(str),
                       "num": rng.integers(0, 4, n),
                       "withna": pd.Series(rng.choice(np.array(["a", "b", None], dtype=object), n)),
                       "cat": pd.Categorical(rng.choice(["lo", "hi", "mid"], n), categories=["lo", "mid", "hi", "unused"])})
    dm = dd.design_matrix(sd, "~ 1 + condition + batch")
    c, b = sd["condition"].to_numpy(dtype=object), sd["batch"].to_numpy(dtype=object)
    want = np.concatenate([np.ones((n, 1)), (c == "1")[:, None],
                           b[:, None] == np.array([str(i) for i in range(1, 8)], dtype=object)[None, :]], axis=1).astype(float)
    np.testing.assert_array_equal(np.asarray(dm), want)
    assert list(dm.design_info.column_names) == ["Intercept", "condition[T.1]"] + ["batch[T.%d]" % i for i in range(1, 8)]
    dm2 = dd.design_matrix(sd, "~ 1 + num + withna")
    assert list(dm2.design_info.column_names) == ["Intercept", "num[T.1]", "num[T.2]", "num[T.3]", "withna[T.b]"]
    x2, w = np.asarray(dm2), sd["withna"].to_numpy(dtype=object)
    np.testing.assert_array_equal(x2[:, 4], (w == "b").astype(float))          # missing values: all-zero row of the factor
    np.testing.assert_array_equal(x2[:, 1], (sd["num"].to_numpy() == 1).astype(float))
    dm3 = dd.design_matrix(sd, "~ 1 + cat")          # an existing categorical keeps its level order; unused levels drop out
    assert list(dm3.design_info.column_names) == ["Intercept", "cat[T.mid]", "cat[T.hi]"]
