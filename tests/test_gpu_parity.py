"""GPU (-m gpu): parity of the CUDA path with the oracle and the golden fixtures, called through the C ABI
(ctypes) and through the reference-shaped API.

Tolerances (BASELINE.json north_star): <= 1e-5 relative on converged coefficients (absolute floor 2e-6 for
coefficients near zero), <= 1e-4 relative on log p-values (where the reference's 1 - cdf form still carries
digits), bit-exact rank sums / U statistics and convergence flags / niter (against oracle method_b="newton",
the restatement of the algorithm the kernels run)."""
import os

import numpy as np
import pandas as pd
import pytest
import scipy.sparse
import scipy.stats

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import oracle.nb_glm as onb          # noqa: E402
import oracle.stats as ost           # noqa: E402


@pytest.fixture(scope="module")
def de():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffxpy_b200.api as de_
    return de_


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stats_golden.npz"))


COEF_RTOL, COEF_ATOL = 1e-5, 2e-6


# ---------------------------------------------------------------------------------------------------
# K_stats against the reference's own stats.py outputs
# ---------------------------------------------------------------------------------------------------
def test_wald_test_kernel(de, gold):
    from diffxpy_b200.stats import stats as dstats
    p, lp = dstats.wald_test_full(gold["wald_theta"], gold["wald_sd"])
    np.testing.assert_allclose(p, gold["wald_p"], rtol=0, atol=4e-16)
    assert p[2] == 0.0 and gold["wald_p"][2] == 0.0            # saturation of 1 - cdf kept
    z = np.abs(gold["wald_theta"] / np.maximum(gold["wald_sd"], np.nextafter(0, 1)))
    ok = np.isfinite(z)
    ref_lp = np.log(2) + scipy.stats.norm.logsf(z[ok])
    np.testing.assert_allclose(lp[ok], ref_lp, rtol=1e-12, atol=1e-14)
    carries = gold["wald_p"] > 1e-11
    np.testing.assert_allclose(lp[carries], np.log(gold["wald_p"][carries]), rtol=1e-4, atol=1e-12)


def test_wald_nan_sd_is_nan_not_zero(de, gold):
    """Singular Fisher information -> NaN sd -> p = NaN (not 0), excluded from the q-values (reference golden)."""
    from diffxpy_b200.stats import stats as dstats
    from diffxpy_b200.testing import correction
    p, lp = dstats.wald_test_full(gold["waldnan_theta"], gold["waldnan_sd"])
    np.testing.assert_allclose(p, gold["waldnan_p"], rtol=0, atol=4e-16, equal_nan=True)
    assert np.isnan(p[0]) and np.isnan(p[2]) and np.isnan(p[3]) and np.isnan(lp[0])
    q = correction.correct(p)
    np.testing.assert_array_equal(q, ost.bh_correct(gold["waldnan_p"]))
    # the same rule inside dx_wald_from_fit: a NaN diagonal of fisher_inv (failed final factorisation) -> NaN row
    import torch
    from diffxpy_b200 import _lib
    from diffxpy_b200.device import ptr, stream_ptr
    dev = torch.device("cuda")
    a = torch.tensor([[0.1, 0.2, 0.3], [1.0, -2.0, 0.5]], dtype=torch.float64, device=dev)        # [p][G]
    fi = torch.eye(2, dtype=torch.float64, device=dev).repeat(3, 1, 1).contiguous()
    fi[1] = float("nan")
    out = torch.empty((4, 3), dtype=torch.float64, device=dev)
    _lib.call("dx_wald_from_fit", 3, 2, 1, ptr(a), ptr(fi), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]), stream_ptr())
    o = out.cpu().numpy()
    assert np.isnan(o[2, 1]) and np.isnan(o[3, 1]) and np.isnan(o[1, 1])
    assert np.all(np.isfinite(o[2, [0, 2]]))


def test_wald_chisq_kernel(de, gold):
    p = de.stats.wald_test_chisq(gold["chisq_theta"], gold["chisq_cov"])
    assert np.isnan(p[7])
    ok = ~np.isnan(gold["chisq_p"])
    np.testing.assert_allclose(p[ok], gold["chisq_p"][ok], rtol=1e-9, atol=1e-15)


def test_lrt_kernel(de, gold):
    p = de.stats.likelihood_ratio_test(gold["lrt_ll_full"], gold["lrt_ll_red"], int(gold["lrt_df"][0]), int(gold["lrt_df"][1]))
    np.testing.assert_allclose(p, gold["lrt_p"], rtol=1e-10, atol=1e-15)


def test_t_test_moments_kernel(de, gold):
    p = de.stats.t_test_moments(gold["tt_mu0"], gold["tt_mu1"], gold["tt_var0"], gold["tt_var1"],
                                int(gold["tt_n"][0]), int(gold["tt_n"][1]))
    np.testing.assert_allclose(p, gold["tt_p"], rtol=1e-9, atol=1e-300)


def test_two_coef_z_kernel(de, gold):
    th, sd = gold["wald_theta"], gold["wald_sd"]
    p = de.stats.two_coef_z_test(th, th[::-1].copy(), sd + 0.1, sd[::-1].copy() + 0.1)
    np.testing.assert_allclose(p, gold["z_p"], rtol=0, atol=1e-15)


def test_bh_on_device(de, gold):
    from diffxpy_b200.testing import correction
    q = correction.correct(gold["bh_p"])
    np.testing.assert_allclose(q, gold["bh_q"], rtol=1e-14, atol=0)
    p = gold["bh_p"].copy()
    p[[3, 50]] = np.nan
    np.testing.assert_allclose(correction.correct(p), ost.bh_correct(p), rtol=1e-14, atol=0, equal_nan=True)
    # heavy ties (rank-test p-values repeat), exact zeros / ones, NaNs: the sort-free kernel and the sort path agree bit for bit
    rng = np.random.default_rng(5)
    p = np.round(rng.uniform(size=5000), 2)
    p[::97] = np.nan
    p[1::101] = 0.0
    want = ost.bh_correct(p)
    np.testing.assert_array_equal(correction.correct(p), want)
    old = correction.BH_KERNEL_MAX
    try:
        correction.BH_KERNEL_MAX = 0          # force the torch.sort path
        np.testing.assert_array_equal(correction.correct(p), want)
    finally:
        correction.BH_KERNEL_MAX = old
    assert np.all(np.isnan(correction.correct(np.full(7, np.nan))))
    # the bucketed rank count: p-values over 300 decades, denormals, exact 0 / 1, one value repeated thousands of times
    rng = np.random.default_rng(6)
    p = np.concatenate([rng.uniform(size=30000), 10.0 ** rng.uniform(-320, 0, 20000), np.full(3000, 1.0), np.zeros(50),
                        np.full(2500, 0.25), np.full(40, np.nan), [2.0 ** -16, np.nextafter(2.0 ** -16, 0), 5e-324]])
    rng.shuffle(p)
    np.testing.assert_array_equal(correction.correct(p), ost.bh_correct(p))
    np.testing.assert_array_equal(correction.correct(p[:1]), ost.bh_correct(p[:1]))


# ---------------------------------------------------------------------------------------------------
# K_ingest: bit-exact sufficient statistics
# ---------------------------------------------------------------------------------------------------
def _ingest(x, grp, k):
    from diffxpy_b200.device import CountShard
    shard = CountShard.from_host(x)
    cg = torch.from_numpy(np.ascontiguousarray(grp, dtype=np.uint8)).cuda()
    suff = shard.group_suffstats(cg, k)
    torch.cuda.synchronize()
    return shard, suff


@pytest.mark.parametrize("k", [1, 2, 16, 200])
def test_ingest_tail_counts_bit_exact(de, k):
    rng = np.random.default_rng(k)
    n, g = 3000, 40
    lam = np.exp(rng.uniform(np.log(0.02), np.log(30), g))
    x = rng.poisson(rng.gamma(0.8, lam / 0.8, size=(n, g))).astype(np.float32)
    x[:, 0] = 0
    x[::7, 1] = 2.5            # non-integer -> overflow
    x[::11, 2] = -1.0          # negative -> overflow
    grp = rng.integers(0, k, n).astype(np.uint8)
    grp[5:9] = 255             # ignored cells
    shard, suff = _ingest(x, grp, k)
    tail = suff.tail.cpu().numpy().view(np.uint32)
    oc = suff.ovf_count.cpu().numpy()
    indptr = shard.indptr.cpu().numpy()
    ov, og = suff.ovf_val.cpu().numpy(), suff.ovf_group.cpu().numpy()
    for gi in range(g):
        col = x[:, gi]
        for kk in range(k):
            v = col[grp == kk]
            isbin = (v >= 1) & (v <= 64) & (v == np.floor(v))
            want = np.array([(v[isbin] > j).sum() for j in range(64)])
            np.testing.assert_array_equal(tail[gi, kk], want)
        keep = grp != 255
        v = col[keep]
        gk = grp[keep]
        isovf = (v != 0) & ~((v >= 1) & (v <= 64) & (v == np.floor(v)))
        assert oc[gi] == isovf.sum()
        got = sorted(zip(ov[indptr[gi]:indptr[gi] + oc[gi]].tolist(), og[indptr[gi]:indptr[gi] + oc[gi]].tolist()))
        assert got == sorted(zip(v[isovf].tolist(), gk[isovf].tolist()))


def test_ingest_ex_order_and_integer_flag(de):
    """dx_group_suffstats_ex (largest-first gene order, integer-count fast path) == dx_group_suffstats."""
    from diffxpy_b200.device import CountShard, ptr, stream_ptr
    from diffxpy_b200 import _lib
    rng = np.random.default_rng(11)
    n, g, k = 5000, 300, 16
    lam = np.exp(rng.uniform(np.log(0.01), np.log(20), g))
    x = rng.poisson(rng.gamma(0.7, lam / 0.7, size=(n, g))).astype(np.float32)
    x[:, 5] = 0
    x[:, 7] = 70                     # above the direct bins -> overflow entries even for integer data
    grp = rng.integers(0, k, n).astype(np.uint8)
    shard = CountShard.from_host(x)
    assert shard.integer_counts
    order = shard.gene_order.cpu().numpy()
    assert sorted(order.tolist()) == list(range(g))
    nnz_g = np.diff(shard.indptr.cpu().numpy())
    assert np.all(np.diff(nnz_g[order]) <= 0)
    cg = torch.from_numpy(grp).cuda()
    suff = shard.group_suffstats(cg, k)                     # _ex path
    tail2 = torch.empty_like(suff.tail); oc2 = torch.empty_like(suff.ovf_count)
    ov2 = torch.zeros_like(suff.ovf_val); og2 = torch.zeros_like(suff.ovf_group)
    _lib.call("dx_group_suffstats", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cg), k, ptr(tail2), ptr(oc2),
              ptr(ov2), ptr(og2), stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(suff.tail, tail2) and torch.equal(suff.ovf_count, oc2)
    assert int(oc2[7].item()) == n
    x[3, 9] = 0.5
    assert not CountShard.from_host(x).integer_counts


def test_ingest_long_ragged_genes_all_alignments(de):
    """TMA ingest edge cases: empty genes (first / middle / last), every start alignment modulo 4 elements, genes longer
    than one counter-flush interval (> 255 entries per lane), exact multiples of the tile size, K = 2 and K = 32."""
    from diffxpy_b200.device import CountShard
    rng = np.random.default_rng(3)
    n = 40000
    lens = [0, 1, 2, 3, 5, 255, 256, 257, 511, 512, 513, 0, 0, 1023, 1024, 8160, 8161, 8192, 30000, 39999, 7, 0]
    cols, rws, vls = [], [], []
    for gi, m in enumerate(lens):
        r = np.sort(rng.choice(n, size=m, replace=False))
        v = rng.integers(1, 12, size=m).astype(np.float32)
        if m > 100:
            v[::97] = 70.0          # above the direct bins
            v[5::89] = 2.5          # non-integer
        cols.append(np.full(m, gi)); rws.append(r); vls.append(v)
    x = scipy.sparse.csc_matrix((np.concatenate(vls), (np.concatenate(rws), np.concatenate(cols))), shape=(n, len(lens)))
    dense = np.asarray(x.todense())
    for k in (2, 32):
        grp = rng.integers(0, k, n).astype(np.uint8)
        grp[::501] = 255
        shard = CountShard.from_host(x)
        assert not shard.integer_counts
        suff = shard.group_suffstats(torch.from_numpy(grp).cuda(), k)
        torch.cuda.synchronize()
        tail = suff.tail.cpu().numpy().view(np.uint32)
        oc = suff.ovf_count.cpu().numpy()
        for gi in range(len(lens)):
            col = dense[:, gi]
            isbin = (col >= 1) & (col <= 64) & (col == np.floor(col))
            for kk in range(k):
                sel = isbin & (grp == kk)
                want = np.array([(col[sel] > j).sum() for j in range(64)])
                np.testing.assert_array_equal(tail[gi, kk], want, err_msg="gene %d group %d" % (gi, kk))
            assert oc[gi] == ((col != 0) & ~isbin & (grp != 255)).sum()


def test_ingest_csr_csc_dense_inputs_agree(de):
    rng = np.random.default_rng(0)
    x = rng.poisson(0.7, size=(500, 30)).astype(np.float32)
    grp = rng.integers(0, 3, 500)
    outs = []
    for data in (x, scipy.sparse.csr_matrix(x), scipy.sparse.csc_matrix(x)):
        _, suff = _ingest(data, grp, 3)
        outs.append(suff.tail.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])


@pytest.mark.parametrize("shape,density", [((3000, 700), 0.12), ((257, 40000), 0.002), ((1, 5), 1.0), ((4000, 3), 0.5),
                                           ((20000, 300), 0.6), ((700, 2100), 0.9)])
def test_k8_csr_to_csc_bit_exact(de, shape, density):
    """dx_csr_to_csc against scipy's tocsc(): same pointers, cells ascending inside every gene, same values -- whole matrix
    and gene ranges (what a gene-sharded run keeps), with empty cells and empty genes."""
    from diffxpy_b200.device import CountShard
    rng = np.random.default_rng(shape[0] + shape[1])
    x = scipy.sparse.random(shape[0], shape[1], density=density, format="csr", random_state=rng,
                            data_rvs=lambda k: rng.integers(1, 90, k).astype(np.float32)).astype(np.float32)
    x.sort_indices()
    if shape[0] > 10:
        x = scipy.sparse.vstack([x[:5], scipy.sparse.csr_matrix((3, shape[1]), dtype=np.float32), x[5:]]).tocsr()   # empty cells
    want = x.tocsc()
    want.sort_indices()
    ranges = [None, (0, shape[1] // 2), (shape[1] // 3, shape[1])] if shape[1] > 3 else [None]
    for gr in ranges:
        for staged in ("1", "0"):          # dx_csr_to_csc_staged (default) and dx_csr_to_csc
            os.environ["DXB200_K8_STAGED"] = staged
            try:
                shard = CountShard.from_host(x, gene_range=gr)
            finally:
                del os.environ["DXB200_K8_STAGED"]
            lo, hi = (0, shape[1]) if gr is None else gr
            sub = want[:, lo:hi]
            np.testing.assert_array_equal(shard.indptr.cpu().numpy(), sub.indptr.astype(np.int64))
            np.testing.assert_array_equal(shard.rows.cpu().numpy()[:sub.nnz], sub.indices)
            np.testing.assert_array_equal(shard.vals.cpu().numpy()[:sub.nnz], sub.data)
            assert shard.gene_offset == lo and shard.n_cells == x.shape[0]


# ---------------------------------------------------------------------------------------------------
# K_fit (grouped) against the oracle and the committed goldens
# ---------------------------------------------------------------------------------------------------
def _fit(de, y, xd, zd, **kw):
    from diffxpy_b200.testing.tests import _fit as fit
    g = y.shape[1]
    return fit("nb", y, pd.DataFrame(xd, columns=["c%d" % i for i in range(xd.shape[1])]),
               pd.DataFrame(zd, columns=["s%d" % i for i in range(zd.shape[1])]),
               gene_names=["g%d" % i for i in range(g)], **kw)


def _check_fit(est, ref, flags_exact=True, skip=None):
    keep = np.ones(ref["a_var"].shape[1], dtype=bool) if skip is None else ~skip
    np.testing.assert_allclose(est.a_var[:, keep], ref["a_var"][:, keep], rtol=COEF_RTOL, atol=COEF_ATOL)
    fi_ref = ref["fisher_inv"][keep]
    scale = np.abs(fi_ref).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(est.fisher_inv[keep] - fi_ref) <= 2e-5 * scale + 1e-300), "fisher_inv differs from the oracle"
    np.testing.assert_allclose(est.log_likelihood[keep], ref["log_likelihood"][keep], rtol=1e-9)
    if flags_exact:
        np.testing.assert_array_equal(est.niter, ref["niter"])
        np.testing.assert_array_equal(est.error_codes, ref["error_codes"])


def test_cfg1_golden_and_oracle(de, golden_dir):
    """BASELINE.json configs[0]: 2000 x 100, '~ 1 + condition' (saturated: closed-form location, scale only)."""
    gd = np.load(os.path.join(golden_dir, "nbglm_cfg1.npz"))
    y = gd["x"].astype(np.float64)
    est = _fit(de, y, gd["design_loc"], gd["design_scale"])
    newton = {k[7:]: gd[k] for k in gd.files if k.startswith("newton_")}
    brent = {k[6:]: gd[k] for k in gd.files if k.startswith("brent_")}
    _check_fit(est, newton, flags_exact=True)
    _check_fit(est, brent, flags_exact=False)          # reference schedule with scipy brent
    np.testing.assert_allclose(est.b_var, newton["b_var"], rtol=1e-5, atol=2e-6)
    # the residual gradient is whatever the last scale step left: ~1e-13 when that step was accepted, ~1e-6 when it was
    # reverted (a decision taken on a log-likelihood difference at rounding level)
    np.testing.assert_allclose(est.jacobian, newton["jacobian"], rtol=1e-3, atol=1e-5)


def test_batch_design_golden(de, golden_dir):
    """'~ 1 + condition + batch', low counts; includes genes whose moment init sits on the Poisson plateau."""
    gd = np.load(os.path.join(golden_dir, "nbglm_batch.npz"))
    y = gd["x"].astype(np.float64)
    est = _fit(de, y, gd["design_loc"], gd["design_scale"])
    newton = {k[7:]: gd[k] for k in gd.files if k.startswith("newton_")}
    brent = {k[6:]: gd[k] for k in gd.files if k.startswith("brent_")}
    _check_fit(est, newton, flags_exact=True)
    frozen = (newton["error_codes"] & onb.FLAG_SCALE_FROZEN) != 0
    assert frozen.sum() > 0
    # reference-faithful oracle: compare where the reference arithmetic is defined (off the plateau)
    keep = ~frozen
    np.testing.assert_allclose(est.a_var[:, keep], brent["a_var"][:, keep], rtol=COEF_RTOL, atol=COEF_ATOL)
    np.testing.assert_array_equal(est.niter[keep], brent["niter"][keep])


@pytest.mark.parametrize("fmt", ["dense", "csr"])
def test_fit_size_factor_groups_and_scale_model(de, fmt):
    """Scale model '~ 1 + batch' (several scale coefficients, joint Newton) and group-constant size factors."""
    rng = np.random.default_rng(11)
    n, g = 1200, 24
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 3, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 3)[None])], axis=1).astype(float)
    zd = np.concatenate([np.ones((n, 1)), (batch[:, None] == np.arange(1, 3)[None])], axis=1).astype(float)
    sf = np.array([0.6, 1.0, 1.7])[batch]
    mu = np.exp(rng.uniform(np.log(0.5), np.log(20), g))[None] * np.exp(0.5 * cond)[:, None] * sf[:, None]
    r = np.exp(rng.uniform(-0.5, 1.0, (3, g)))[batch]
    y = rng.poisson(rng.gamma(r, mu / r)).astype(np.float64)
    data = y if fmt == "dense" else scipy.sparse.csr_matrix(y)
    est = _fit(de, data, xd, zd, size_factors=sf)
    ref = onb.fit_nb(y, xd, zd, size_factors=sf, method_b="newton")
    _check_fit(est, ref, flags_exact=True)
    np.testing.assert_allclose(est.b_var, ref["b_var"], rtol=1e-5, atol=1e-6)


def test_cfg2_full_cell_count_against_oracle(de):
    """BASELINE configs[1] at its full cell count (100 000 cells, '~ 1 + condition + batch' with 8 batches, sparse input),
    on a handful of genes the CPU oracle can afford: coefficients, Fisher inverse, ll, niter and flags."""
    rng = np.random.default_rng(2024)
    n, g = 100_000, 10
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 8, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 8)[None])], axis=1).astype(float)
    log_mu = rng.uniform(np.log(0.005), np.log(1.0), g)
    log_mu[0], log_mu[1] = np.log(0.002), np.log(3.0)
    r = rng.uniform(0.5, 4.0, g)
    beta_c = np.where(np.arange(g) % 3 == 0, rng.normal(0, 0.5, g), 0.0)
    beta_b = rng.normal(0, 0.2, (8, g)); beta_b[0] = 0
    mu = np.exp(log_mu[None] + cond[:, None] * beta_c[None] + beta_b[batch])
    y = rng.poisson(rng.gamma(r[None], mu / r[None])).astype(np.float64)
    est = _fit(de, scipy.sparse.csr_matrix(y.astype(np.float32)), xd, np.ones((n, 1)))
    assert est._plan["kind"] == "grouped" and est._plan["k"] == 16
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
    # the schedule's decisions are identical ...
    np.testing.assert_array_equal(est.niter, ref["niter"])
    np.testing.assert_array_equal(est.error_codes, ref["error_codes"])
    np.testing.assert_allclose(est.log_likelihood, ref["log_likelihood"], rtol=1e-11)
    # ... and so are the coefficients, except where the LAST accepted / reverted step improves the likelihood by less
    # than its rounding noise (|ll| ~ 1e5, noise ~ 1e-8): batchglm's rule "revert the step if ll got worse" then flips
    # between the two summation orders, and the coefficient moves by that last step, ~ sqrt(noise / information)
    # ~ 1e-6 absolute, i.e. < 1e-4 standard errors.  Tolerance: 1e-5 relative + 2e-4 standard errors.
    sd_all = np.sqrt(np.diagonal(ref["fisher_inv"], axis1=-2, axis2=-1).T)
    tol = 1e-5 * np.abs(ref["a_var"]) + 2e-6 + 2e-4 * sd_all
    assert np.all(np.abs(est.a_var - ref["a_var"]) <= tol), np.max(np.abs(est.a_var - ref["a_var"]) / sd_all)
    exact = np.max(np.abs(est.a_var - ref["a_var"]), axis=0) < 1e-9
    assert exact.sum() >= g // 2, "most genes must agree to rounding"
    np.testing.assert_allclose(est.b_var, ref["b_var"], rtol=1e-4, atol=2e-5)
    fi_ref = ref["fisher_inv"]
    scale = np.abs(fi_ref).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(est.fisher_inv - fi_ref) <= 2e-5 * scale)
    sd = np.sqrt(np.diagonal(est.fisher_inv, axis1=-2, axis2=-1).T[1])
    ref_sd = np.sqrt(ref["fisher_inv"][:, 1, 1])
    np.testing.assert_allclose(sd, ref_sd, rtol=1e-5)
    p_ref = ost.wald_test(ref["a_var"][1], ref_sd)
    from diffxpy_b200.stats import stats as dstats
    p_gpu = dstats.wald_test(est.a_var[1], sd)
    keep = p_ref > 1e-11            # the reference's 1 - cdf form carries >= 5 digits there
    np.testing.assert_allclose(np.log(p_gpu[keep]), np.log(p_ref[keep]), rtol=1e-4, atol=1e-4)


def test_fit_high_counts_overflow_path(de):
    """Counts far above the 64 direct bins and non-integer values go through the overflow entries."""
    rng = np.random.default_rng(4)
    n, g = 800, 10
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 2, n)
    xd = np.stack([np.ones(n), cond, batch], axis=1).astype(float)
    mu = np.exp(rng.uniform(np.log(30), np.log(3000), g))[None] * np.exp(0.3 * cond)[:, None]
    y = rng.poisson(rng.gamma(2.0, mu / 2.0)).astype(np.float64)
    y[:, 0] = y[:, 0] / 4.0         # non-integer "normalised" values
    est = _fit(de, y, xd, np.ones((n, 1)))
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
    _check_fit(est, ref, flags_exact=True)


@pytest.mark.parametrize("scale_model", ["const", "batch"])
def test_fit_packed_overflow_lists(de, scale_model):
    """Highly expressed genes: K_pack rewrites their overflow lists as (value, multiplicity) records.  The packed fit
    must agree with the oracle like the plain one, with the plain fit to summation order, and keep entries that cannot
    be packed (non-integer values, counts beyond the table) verbatim."""
    from diffxpy_b200 import _lib
    from diffxpy_b200.device import ptr, stream_ptr
    rng = np.random.default_rng(17)
    n, g = 20_000, 12
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 4, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 4)[None])], axis=1).astype(float)
    zd = np.ones((n, 1)) if scale_model == "const" else np.delete(xd, 1, axis=1)
    mu = np.exp(rng.uniform(np.log(20), np.log(400), g))[None] * np.exp(0.3 * cond)[:, None] * np.exp(0.1 * batch)[:, None]
    mu[:, 2] *= 0.02                 # a gene with a short overflow list (left alone)
    y = rng.poisson(rng.gamma(1.5, mu / 1.5)).astype(np.float64)
    y[:40, 0] += 0.5                 # a few entries that must stay verbatim
    y[:5, 3] = 70000.0               # beyond the table range
    y[:, 1] = y[:, 1] / 4.0          # nothing to pack: (almost) every value is distinct / non-integer
    est = _fit(de, scipy.sparse.csc_matrix(y.astype(np.float32)), xd, zd)
    suff = est._plan["suff"]
    packed = suff.ovf_packed.cpu().numpy()
    oc = suff.ovf_count.cpu().numpy()
    assert packed[0, 0] > 0 and packed[3, 0] > 0 and (packed[4:, 0] > 0).all(), packed
    assert packed[1, 0] == 0 and packed[2, 0] == 0
    assert (packed[:, 1] <= packed[:, 0]).all() and (16 * 8 + 2 * packed.sum(axis=1) <= np.maximum(oc, 16 * 8)).all()
    assert packed[5, 1] * 8 < oc[5], "pooled records must be far fewer than entries"
    ref = onb.fit_nb(y, xd, zd, method_b="newton")
    _check_fit(est, ref, flags_exact=True)
    np.testing.assert_allclose(est.b_var, ref["b_var"], rtol=1e-5, atol=2e-6)
    # plain lists through the C ABI on a fresh K_ingest pass: same decisions, same likelihood to rounding
    plan, shard = est._plan, est.shard
    plain = shard.group_suffstats(plan["cell_group"], plan["k"])
    p, ps = xd.shape[1], zd.shape[1]
    dev = plain.tail.device
    f64 = dict(dtype=torch.float64, device=dev)
    a2, b2 = torch.zeros((p, g), **f64), torch.zeros((ps, g), **f64)
    fi2, ll2, jac2 = torch.zeros((g, p, p), **f64), torch.zeros(g, **f64), torch.zeros(g, **f64)
    ni2, fl2 = torch.zeros(g, dtype=torch.int32, device=dev), torch.zeros(g, dtype=torch.int32, device=dev)
    opts = _lib.DxFitOpts(max_steps=1000, update_b_freq=5, train_loc=1, train_scale=1, init_a=_lib.DX_INIT_CLOSED_FORM,
                          init_b=_lib.DX_INIT_STANDARD, newton_max_iter=100, reserved=0, lltol=1e-10, newton_xtol=1e-8)
    _lib.call("dx_nb_fit_grouped", g, plan["k"], p, ps, ptr(plan["xu_loc"]), ptr(plan["xu_scale"]), ptr(plan["offset"]),
              ptr(plan["n_k"]), ptr(plan["pinv"]), ptr(plan["loc_group"]), int(plan["kl"]), ptr(plain.tail),
              ptr(plain.ovf_count), ptr(shard.indptr), ptr(plain.ovf_val), ptr(plain.ovf_group), opts, ptr(a2), ptr(b2),
              ptr(fi2), ptr(ll2), ptr(jac2), ptr(ni2), ptr(fl2), stream_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ni2.cpu().numpy(), est.niter)
    np.testing.assert_allclose(ll2.cpu().numpy(), est.log_likelihood, rtol=1e-12)
    np.testing.assert_allclose(a2.cpu().numpy(), est.a_var, rtol=1e-8, atol=1e-9)


def test_pack_overflow_records_bit_exact(de):
    """K_pack: unpacking the (value, multiplicity) records gives back the plain overflow list as a multiset, per group
    and pooled; the per-group header holds the sums K_fit needs at set-up; short lists are left untouched."""
    from scipy.special import gammaln
    rng = np.random.default_rng(23)
    n, g, k = 30_000, 16, 5
    grp = rng.integers(0, k, n)
    mu = np.exp(rng.uniform(np.log(0.5), np.log(500), g))
    y = rng.poisson(rng.gamma(1.2, mu[None] / 1.2, size=(n, g))).astype(np.float32)
    y[:17, 0] += 0.25                                        # kept verbatim
    y[:3, 1] = 50000.0                                       # beyond the table
    shard, plain = _ingest(y, grp, k)
    _, pk = _ingest(y, grp, k)
    pk.pack_overflow()
    torch.cuda.synchronize()
    oc, ip = plain.ovf_count.cpu().numpy(), shard.indptr.cpu().numpy()
    pv, pg = plain.ovf_val.cpu().numpy(), plain.ovf_group.cpu().numpy()
    qv, qg = pk.ovf_val.cpu().numpy(), pk.ovf_group.cpu().numpy()
    packed = pk.ovf_packed.cpu().numpy()
    assert (packed[:, 0] > 0).sum() >= 4 and (packed[:, 0] == 0).sum() >= 2
    np.testing.assert_array_equal(pk.ovf_count.cpu().numpy(), oc)
    for gi in range(g):
        na, nb = packed[gi]
        s = ip[gi]
        if na == 0:
            np.testing.assert_array_equal(pv[s:s + oc[gi]], qv[s:s + oc[gi]])
            np.testing.assert_array_equal(pg[s:s + oc[gi]], qg[s:s + oc[gi]])
            continue
        assert 8 * k + 2 * na + 2 * nb <= oc[gi]
        vals, grps = pv[s:s + oc[gi]].astype(np.float64), pg[s:s + oc[gi]]
        head = np.frombuffer(qv[s:s + 8 * k].tobytes(), dtype=np.float64).reshape(k, 4)
        rec_a = qv[s + 8 * k: s + 8 * k + 2 * na]
        v_a, m_a = rec_a[0::2].astype(np.float64), rec_a[1::2].copy().view(np.uint32).astype(np.int64)
        g_a = qg[s + 8 * k: s + 8 * k + 2 * na][0::2]
        rec_b = qv[s + 8 * k + 2 * na: s + 8 * k + 2 * na + 2 * nb]
        v_b, m_b = rec_b[0::2].astype(np.float64), rec_b[1::2].copy().view(np.uint32).astype(np.int64)
        for kk in range(k):
            want = np.sort(vals[grps == kk])
            np.testing.assert_array_equal(np.sort(np.repeat(v_a[g_a == kk], m_a[g_a == kk])), want)
            np.testing.assert_allclose(head[kk], [want.sum(), (want ** 2).sum(), gammaln(want + 1).sum(), len(want)], rtol=1e-12)
        np.testing.assert_array_equal(np.sort(np.repeat(v_b, m_b)), np.sort(vals))
        # records of the table part are ordered (group, value); verbatim entries follow in their original order
        tab = m_a > 1
        key = g_a[tab].astype(np.int64) * 100000 + v_a[tab].astype(np.int64)
        assert np.all(np.diff(key) > 0)
    from diffxpy_b200.stats import stats as dstats
    with pytest.raises(ValueError):
        dstats.group_pair_device(pk, 0, 1, 10, 10)


def test_quick_scale_and_given_init(de):
    rng = np.random.default_rng(8)
    n, g = 900, 12
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 3, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 3)[None])], axis=1).astype(float)
    mu = np.exp(rng.uniform(np.log(1), np.log(50), g))[None] * np.exp(0.3 * cond)[:, None]
    y = rng.poisson(rng.gamma(1.5, mu / 1.5)).astype(np.float64)
    est = _fit(de, y, xd, np.ones((n, 1)), quick_scale=True)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), quick_scale=True, method_b="newton")
    _check_fit(est, ref, flags_exact=True)
    a0 = ref["a_var"] + 0.05
    b0 = ref["b_var"] * 0 + 0.3
    est2 = _fit(de, y, xd, np.ones((n, 1)), init_a=a0, init_b=b0)
    ref2 = onb.fit_nb(y, xd, np.ones((n, 1)), init_a=a0, init_b=b0, method_b="newton")
    _check_fit(est2, ref2, flags_exact=True)


def test_all_zero_and_single_count_genes_do_not_crash(de):
    rng = np.random.default_rng(2)
    n = 600
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 2, n)
    xd = np.stack([np.ones(n), cond, batch], axis=1).astype(float)
    y = rng.poisson(1.0, size=(n, 5)).astype(np.float64)
    y[:, 0] = 0
    y[:, 1] = 0
    y[17, 1] = 1
    est = _fit(de, y, xd, np.ones((n, 1)))
    assert est.error_codes[0] & 16
    assert np.all(np.isfinite(est.a_var[:, 2:]))
    ref = onb.fit_nb(y[:, 2:], xd, np.ones((n, 1)), method_b="newton")
    np.testing.assert_allclose(est.a_var[:, 2:], ref["a_var"], rtol=COEF_RTOL, atol=COEF_ATOL)


# ---------------------------------------------------------------------------------------------------
# de.test.wald / lrt through the reference-shaped API
# ---------------------------------------------------------------------------------------------------
def _sim_counts(seed, n, g, n_batch=4, lo=0.2, hi=20.0, frac_de=0.0):
    rng = np.random.default_rng(seed)
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, n_batch, n)
    beta = np.where(np.arange(g) < int(g * frac_de), rng.normal(0, 0.8, g), 0.0)
    bb = rng.normal(0, 0.2, (n_batch, g))
    bb[0] = 0
    mu = np.exp(rng.uniform(np.log(lo), np.log(hi), g)[None] + cond[:, None] * beta[None] + bb[batch])
    r = rng.uniform(0.7, 4, g)
    y = rng.poisson(rng.gamma(r[None], mu / r[None])).astype(np.float64)
    sd = pd.DataFrame({"condition": cond.astype(str), "batch": batch.astype(str)})
    return y, sd, beta != 0


def test_wald_summary_columns_pvalues_and_logp(de):
    y, sd, _ = _sim_counts(21, 1000, 60, frac_de=0.5)
    names = ["gene%d" % i for i in range(y.shape[1])]
    res = de.test.wald(data=scipy.sparse.csr_matrix(y), formula_loc="~ 1 + condition + batch",
                       factor_loc_totest="condition", sample_description=sd, gene_names=names)
    summ = res.summary()
    assert list(summ.columns) == ["gene", "pval", "qval", "log2fc", "mean", "zero_mean", "grad", "coef_mle",
                                  "coef_sd", "ll", "err", "niter"]
    from diffxpy_b200.data import design
    xd = np.asarray(design.design_matrix(sd, "~ 1 + condition + batch"))
    ref = onb.fit_nb(y, xd, np.ones((y.shape[0], 1)), method_b="brent")
    ref_sd = np.sqrt(ref["fisher_inv"][:, 1, 1])
    ref_p = ost.wald_test(ref["a_var"][1], ref_sd)
    np.testing.assert_allclose(summ["coef_mle"].values, ref["a_var"][1], rtol=COEF_RTOL, atol=COEF_ATOL)
    np.testing.assert_allclose(summ["coef_sd"].values, ref_sd, rtol=1e-5)
    carries = ref_p > 1e-11
    np.testing.assert_allclose(np.log(summ["pval"].values[carries]), np.log(ref_p[carries]), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res.log_pval[carries], np.log(ref_p[carries]), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(summ["qval"].values, ost.bh_correct(summ["pval"].values), rtol=1e-13)
    np.testing.assert_allclose(summ["mean"].values, y.mean(axis=0), rtol=1e-12)
    np.testing.assert_array_equal(summ["log2fc"].values, summ["coef_mle"].values)     # det.py:742-765 quirk kept
    # wald_repeated on the retained fit: multi-coefficient chi^2 test of the batch factor
    rep = de.test.wald_repeated(res, factor_loc_totest="batch")
    cov = ref["fisher_inv"][:, 2:5, 2:5]
    p_chi = ost.wald_test_chisq(ref["a_var"][2:5], cov)
    np.testing.assert_allclose(rep.pval, p_chi, rtol=2e-4, atol=1e-12)


def test_dense_and_sparse_inputs_identical(de):
    # unit_test/test_data_types.py:66-75
    y, sd, _ = _sim_counts(5, 600, 20)
    names = ["g%d" % i for i in range(y.shape[1])]
    kw = dict(formula_loc="~ 1 + condition + batch", factor_loc_totest="condition", sample_description=sd, gene_names=names)
    p_dense = de.test.wald(data=y, **kw).pval
    p_csr = de.test.wald(data=scipy.sparse.csr_matrix(y), **kw).pval
    p_csc = de.test.wald(data=scipy.sparse.csc_matrix(y), **kw).pval
    np.testing.assert_array_equal(p_dense, p_csr)
    np.testing.assert_array_equal(p_dense, p_csc)


def test_wald_null_calibration_and_power(de):
    # unit_test/test_single_null.py:12-62 (KS uniformity of null p-values) and test_single_de.py:46-64
    y, sd, is_de = _sim_counts(1, 2000, 200, frac_de=0.0, lo=5, hi=200)
    names = ["g%d" % i for i in range(y.shape[1])]
    res = de.test.wald(data=y, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                       sample_description=sd, gene_names=names)
    ks = scipy.stats.kstest(res.pval, "uniform").pvalue
    assert ks > 0.05, "null p-values are not uniform: KS p = %g" % ks
    y, sd, is_de = _sim_counts(2, 2000, 200, frac_de=0.5, lo=5, hi=200)
    res = de.test.wald(data=y, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                       sample_description=sd, gene_names=names)
    q = res.qval
    assert np.mean(q[~is_de] < 0.05) <= 0.1 and np.mean(q[is_de] < 0.05) >= 0.5


def test_lrt_against_oracle(de):
    y, sd, _ = _sim_counts(9, 1000, 40, frac_de=0.5)
    names = ["g%d" % i for i in range(y.shape[1])]
    res = de.test.lrt(data=y, full_formula_loc="~ 1 + condition + batch", reduced_formula_loc="~ 1 + batch",
                      sample_description=sd, gene_names=names)
    from diffxpy_b200.data import design
    xf = np.asarray(design.design_matrix(sd, "~ 1 + condition + batch"))
    xr = np.asarray(design.design_matrix(sd, "~ 1 + batch"))
    one = np.ones((y.shape[0], 1))
    ff = onb.fit_nb(y, xf, one, method_b="newton")
    fr = onb.fit_nb(y, xr, one, method_b="newton")
    p_ref = ost.likelihood_ratio_test(ff["log_likelihood"], fr["log_likelihood"], 6, 5)
    np.testing.assert_allclose(res.full_estim.log_likelihood, ff["log_likelihood"], rtol=1e-9)
    np.testing.assert_allclose(res.reduced_estim.log_likelihood, fr["log_likelihood"], rtol=1e-9)
    carries = p_ref > 1e-11
    np.testing.assert_allclose(np.log(res.pval[carries]), np.log(p_ref[carries]), rtol=1e-4, atol=1e-5)
    summ = res.summary()
    assert list(summ.columns) == ["gene", "pval", "qval", "log2fc", "mean", "zero_mean", "grad", "grad_red"]
    lfc = res.log_fold_change()
    np.testing.assert_allclose(lfc, ff["a_var"][1], rtol=1e-5, atol=2e-6)


def test_constrained_wald_runs(de):
    # unit_test/test_constrained.py:101-145: batches nested in conditions
    rng = np.random.default_rng(3)
    n, g = 800, 30
    cond = np.repeat([0, 1], n // 2)
    batch = np.repeat([0, 1, 2, 3], n // 4)
    mu = np.exp(rng.uniform(np.log(2), np.log(50), g))[None] * np.ones((n, 1))
    y = rng.poisson(rng.gamma(2.0, mu / 2.0)).astype(np.float64)
    sd = pd.DataFrame({"condition": cond.astype(str), "batch": batch.astype(str)})
    names = ["g%d" % i for i in range(g)]
    with pytest.raises(ValueError):
        de.test.wald(data=y, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                     sample_description=sd, gene_names=names)
    res = de.test.wald(data=y, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                       constraints_loc={"batch": "condition"}, sample_description=sd, gene_names=names)
    assert res.model_estim.a_var.shape[0] == 4 and np.all(np.isfinite(res.pval))
    from diffxpy_b200.data import design
    dm, _, cmat, _ = design.constraint_system_from_star(sample_description=sd, formula="~ 1 + condition + batch",
                                                       constraints={"batch": "condition"})
    ref = onb.fit_nb(y, np.asarray(dm), np.ones((n, 1)), constraints_loc=cmat, method_b="newton")
    np.testing.assert_allclose(res.model_estim.a_var, ref["a_var"], rtol=COEF_RTOL, atol=COEF_ATOL)


# ---------------------------------------------------------------------------------------------------
# model-free tests
# ---------------------------------------------------------------------------------------------------
def test_rank_and_t_against_reference_golden(de, gold):
    grouping = gold["mwu_grouping"]
    x = gold["mwu_x"].astype(np.float32).astype(np.float64)      # the shard stores float32 values
    names = ["g%d" % i for i in range(x.shape[1])]
    for data in (x, scipy.sparse.csr_matrix(x)):
        res = de.test.rank_test(data=data, grouping=grouping, gene_names=names)
        ref = ost.two_group_test(x, grouping, "rank")
        np.testing.assert_array_equal(res.u_statistic[~np.isnan(ref["pval"])], gold["mwu_u1"][~np.isnan(ref["pval"])])
        np.testing.assert_allclose(res.pval, ref["pval"], rtol=1e-12, atol=1e-300, equal_nan=True)
        run = ~np.isnan(ref["pval"])
        np.testing.assert_allclose(res.pval[run], gold["mwu_p_dense"][run], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(res.summary()["log2fc"].values, ref["logfc"] / np.log(2), rtol=1e-12)
        tt = de.test.t_test(data=data, grouping=grouping, gene_names=names)
        ref_t = ost.two_group_test(x, grouping, "t")
        np.testing.assert_allclose(tt.pval, ref_t["pval"], rtol=1e-8, atol=1e-300, equal_nan=True)


def test_rank_and_t_against_scipy(de):
    # unit_test/test_single_external_libs.py:25-112
    rng = np.random.default_rng(1)
    n, g = 2000, 100
    cond = np.arange(n) % 2
    mu = np.exp(rng.uniform(np.log(0.1), np.log(500), g))[None] * np.ones((n, 1))
    x = rng.poisson(rng.gamma(1.5, mu / 1.5)).astype(np.float64)
    names = ["g%d" % i for i in range(g)]
    sd = pd.DataFrame({"condition": cond})
    a, b = x[cond == 0], x[cond == 1]
    tt = de.test.t_test(data=x, grouping="condition", sample_description=sd, gene_names=names)
    p_t = scipy.stats.ttest_ind(a, b, axis=0, equal_var=False).pvalue
    assert np.max(np.abs(tt.pval - p_t)) < 1e-3 and np.max(np.abs(np.log(tt.pval + 1e-200) - np.log(p_t + 1e-200))) < 1e-1
    rk = de.test.rank_test(data=x, grouping="condition", sample_description=sd, gene_names=names)
    mw = [scipy.stats.mannwhitneyu(a[:, i], b[:, i], use_continuity=True, alternative="two-sided") for i in range(g)]
    np.testing.assert_array_equal(rk.u_statistic, np.array([m.statistic for m in mw]))      # bit exact
    np.testing.assert_allclose(rk.pval, np.array([m.pvalue for m in mw]), rtol=1e-12, atol=1e-300)


def test_rank_test_long_non_count_lists(de):
    """Log-normalised data and highly expressed genes: every non-zero is a non-count value, far more of them than the
    on-chip sort of K_twogroup holds -> dx_group_pair_tests_big (sort in a device workspace, parallel tie merge).
    U is bit exact against scipy, p to rounding; heavy ties (log1p of small counts), several sort tiles, padding."""
    rng = np.random.default_rng(7)
    n, g = 40_000, 10
    cond = (rng.uniform(size=n) < 0.4).astype(int)
    mu = np.exp(rng.uniform(np.log(0.05), np.log(200), g))
    mu[0], mu[1], mu[2] = 0.02, 3.0, 400.0
    cnt = rng.poisson(rng.gamma(1.5, mu[None] / 1.5, size=(n, g))).astype(np.float64)
    cnt[:, 3] *= np.exp(0.2 * cond)                          # a shifted gene
    sf = rng.uniform(0.5, 2.0, n)
    x = np.log1p(cnt / sf[:, None])                          # (almost) all non-zeros distinct
    x[:, 4] = np.log1p(cnt[:, 4])                            # heavy ties among non-count values
    x[:, 5] = cnt[:, 5]                                      # plain counts, many above 64
    x[:, 6] = -x[:, 6]                                       # negative values sort below the zero block
    x = x.astype(np.float32).astype(np.float64)              # the shard stores float32: compare on what it holds
    names = ["g%d" % i for i in range(g)]
    sd = pd.DataFrame({"condition": cond})
    a, b = x[cond == 0], x[cond == 1]
    assert ((x != 0).sum(axis=0) > 20_000).sum() >= 5
    for data in (x, scipy.sparse.csr_matrix(x)):
        rk = de.test.rank_test(data=data, grouping="condition", sample_description=sd, gene_names=names)
        mw = [scipy.stats.mannwhitneyu(a[:, i], b[:, i], use_continuity=True, alternative="two-sided", method="asymptotic")
              for i in range(g)]
        np.testing.assert_array_equal(rk.u_statistic, np.array([m.statistic for m in mw]))
        np.testing.assert_allclose(rk.pval, np.array([m.pvalue for m in mw]), rtol=1e-11, atol=1e-300)
    # several groups from one ingest pass: pairwise and versus-rest take the same path
    grp3 = rng.integers(0, 3, n)
    sd3 = pd.DataFrame({"g": grp3.astype(str)})
    pw = de.test.pairwise(data=x, grouping="g", test="rank", lazy=False, sample_description=sd3, gene_names=names)
    p01 = np.array([scipy.stats.mannwhitneyu(x[grp3 == 0][:, i], x[grp3 == 1][:, i], use_continuity=True,
                                             alternative="two-sided", method="asymptotic").pvalue for i in range(g)])
    groups = [str(v) for v in pw.groups]
    i0, i1 = groups.index("0"), groups.index("1")
    np.testing.assert_allclose(pw.pval[i0, i1], p01, rtol=1e-11, atol=1e-300)


def test_extreme_values_rules(de):
    # unit_test/test_extreme_values.py:11-73
    rng = np.random.default_rng(1)
    x = rng.poisson(3.0, size=(1000, 10)).astype(np.float64)
    x[:, 0] = 0
    x[:, 1] = 5
    grouping = rng.integers(0, 2, 1000)
    names = ["g%d" % i for i in range(10)]
    for fn in (de.test.t_test, de.test.rank_test):
        test = fn(data=x, grouping=grouping, gene_names=names, is_sig_zerovar=True)
        assert np.isnan(test.pval[0]) and test.pval[1] == 1
        s = test.summary()
        assert "zero_variance" in s.columns and bool(s["zero_variance"].values[1])


def test_rank_three_groups_uses_first_two(de):
    rng = np.random.default_rng(6)
    x = rng.poisson(2.0, size=(300, 8)).astype(np.float64)
    grouping = np.array(["b", "a", "c"])[rng.integers(0, 3, 300)]
    grouping[0], grouping[1] = "b", "a"
    names = ["g%d" % i for i in range(8)]
    res = de.test.rank_test(data=x, grouping=grouping, gene_names=names)
    ref = ost.two_group_test(x, grouping, "rank")
    np.testing.assert_allclose(res.pval, ref["pval"], rtol=1e-12, equal_nan=True)


def test_host_entry_point_matches_device_path(de, golden_dir):
    """dx_wald_grouped_host (host CSC buffers in, rows out) == the device-pointer path."""
    import ctypes as C
    from diffxpy_b200 import _lib
    gd = np.load(os.path.join(golden_dir, "nbglm_batch.npz"))
    y = gd["x"].astype(np.float64)
    xd = gd["design_loc"]
    est = _fit(de, y, xd, gd["design_scale"])
    from diffxpy_b200.estimator import group_rows
    n, g = y.shape
    p = xd.shape[1]
    uniq, inv = group_rows(np.concatenate([xd, np.ones((n, 1))], axis=1))
    k = uniq.shape[0]
    csc = scipy.sparse.csc_matrix(y.astype(np.float32))
    indptr = np.ascontiguousarray(csc.indptr, dtype=np.int64)
    rows = np.ascontiguousarray(csc.indices, dtype=np.int32)
    vals = np.ascontiguousarray(csc.data, dtype=np.float32)
    grp = inv.astype(np.uint8)
    xu = np.ascontiguousarray(uniq[:, :p])
    xs = np.ascontiguousarray(uniq[:, p:p + 1])
    nk = np.bincount(inv, minlength=k).astype(np.float64)
    pinv = np.ascontiguousarray(np.linalg.pinv(xu))
    opts = _lib.DxFitOpts(max_steps=1000, update_b_freq=5, train_loc=1, train_scale=1, init_a=_lib.DX_INIT_CLOSED_FORM,
                          init_b=_lib.DX_INIT_STANDARD, newton_max_iter=100, reserved=_lib.DX_OPT_SCALE_CONST,      # as the estimator
                          lltol=1e-10, newton_xtol=1e-8)
    a = np.zeros((p, g)); b = np.zeros((1, g)); fi = np.zeros((g, p, p)); ll = np.zeros(g); jac = np.zeros(g)
    ni = np.zeros(g, dtype=np.int32); fl = np.zeros(g, dtype=np.int32); pv = np.zeros(g); lp = np.zeros(g); sdv = np.zeros(g)
    P = lambda arr: arr.ctypes.data_as(C.c_void_p)
    _lib.call("dx_wald_grouped_host", n, g, P(indptr), P(rows), P(vals), P(grp), k, p, 1, P(xu), P(xs), None, P(nk),
              P(pinv), None, k, C.byref(opts), 1, P(a), P(b), P(fi), P(ll), P(jac), P(ni), P(fl), P(pv), P(lp), P(sdv), None)
    np.testing.assert_array_equal(a, est.a_var)
    np.testing.assert_array_equal(ni, est.niter)
    np.testing.assert_array_equal(fl, est.error_codes)
    np.testing.assert_array_equal(fi, est.fisher_inv)
    # the copy/compute pipeline (several gene chunks) gives the same rows as a single chunk
    os.environ["DXB200_HOST_CHUNKS"] = "5"
    try:
        a2 = np.zeros((p, g)); b2 = np.zeros((1, g)); fi2 = np.zeros((g, p, p)); pv2 = np.zeros(g)
        _lib.call("dx_wald_grouped_host", n, g, P(indptr), P(rows), P(vals), P(grp), k, p, 1, P(xu), P(xs), None, P(nk),
                  P(pinv), None, k, C.byref(opts), 1, P(a2), P(b2), P(fi2), P(ll), P(jac), P(ni), P(fl), P(pv2), P(lp), P(sdv), None)
    finally:
        del os.environ["DXB200_HOST_CHUNKS"]
    np.testing.assert_array_equal(a2, a)
    np.testing.assert_array_equal(b2, b)
    np.testing.assert_array_equal(fi2, fi)
    np.testing.assert_array_equal(pv2, pv)
    np.testing.assert_array_equal(ni, est.niter)


# ---------------------------------------------------------------------------------------------------
# K_fit (per cell): designs that do not collapse into groups
# ---------------------------------------------------------------------------------------------------
def test_percell_size_factors(de):
    """unit_test/test_single_sf_null.py: per-cell size factors U(0.5, 1.5) -> the per-cell streaming kernel."""
    rng = np.random.default_rng(17)
    n, g = 1500, 20
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 3, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 3)[None])], axis=1).astype(float)
    sf = rng.uniform(0.5, 1.5, n)
    mu = np.exp(rng.uniform(np.log(0.3), np.log(30), g))[None] * np.exp(0.4 * cond)[:, None] * sf[:, None]
    y = rng.poisson(rng.gamma(2.0, mu / 2.0)).astype(np.float64)
    y[:, 0] = 0
    y[3, 0] = 2
    for data in (y, scipy.sparse.csr_matrix(y)):
        est = _fit(de, data, xd, np.ones((n, 1)), size_factors=sf)
        assert est._plan["kind"] == "percell"
        ref = onb.fit_nb(y, xd, np.ones((n, 1)), size_factors=sf, method_b="newton")
        degenerate = np.arange(g) == 0        # a single non-zero count: separation, Fisher information ~ 1e-126
        _check_fit(est, ref, flags_exact=True, skip=degenerate)
        np.testing.assert_allclose(est.b_var[:, 1:], ref["b_var"][:, 1:], rtol=1e-5, atol=2e-6)


def test_percell_numeric_covariate_and_multi_scale(de):
    """unit_test/test_numeric_covar.py: numeric covariates; scale model with a numeric covariate too."""
    rng = np.random.default_rng(23)
    n, g = 1200, 12
    cond = rng.integers(0, 2, n)
    t = rng.uniform(0, 1, n)
    xd = np.stack([np.ones(n), cond, t, t * t], axis=1).astype(float)
    zd = np.stack([np.ones(n), t], axis=1).astype(float)
    mu = np.exp(rng.uniform(np.log(1), np.log(40), g))[None] * np.exp(0.5 * cond + 0.8 * t)[:, None]
    r = np.exp(0.3 + 0.5 * t)[:, None]
    y = rng.poisson(rng.gamma(r, mu / r)).astype(np.float64)
    est = _fit(de, y, xd, zd)
    assert est._plan["kind"] == "percell" and est.a_var.shape[0] == 4
    ref = onb.fit_nb(y, xd, zd, method_b="newton")
    _check_fit(est, ref, flags_exact=True)
    np.testing.assert_allclose(est.b_var, ref["b_var"], rtol=1e-5, atol=2e-6)


def test_wald_with_size_factors_null_calibration(de):
    # unit_test/test_single_sf_null.py:12-67
    rng = np.random.default_rng(1)
    n, g = 2000, 200
    cond = rng.integers(0, 2, n)
    sf = rng.uniform(0.5, 1.5, n)
    mu = np.exp(rng.uniform(np.log(5), np.log(200), g))[None] * sf[:, None]
    r = rng.uniform(0.7, 4, g)[None]
    y = rng.poisson(rng.gamma(r, mu / r)).astype(np.float64)
    sd = pd.DataFrame({"condition": cond.astype(str)})
    res = de.test.wald(data=y, formula_loc="~ 1 + condition", factor_loc_totest="condition", sample_description=sd,
                       size_factors=sf, gene_names=["g%d" % i for i in range(g)])
    ks = scipy.stats.kstest(res.pval, "uniform").pvalue
    assert ks > 0.05, "null p-values are not uniform: KS p = %g" % ks


def test_cfg3_design_full_cell_count_eight_dispersion_coefficients(de):
    """BASELINE.json configs[2] at the full cell count: '~1+condition+batch' location (9), '~1+batch' scale model
    (8 dispersion coefficients -> general grouped K_fit, Newton in 8 dimensions), 100k cells, a few genes."""
    rng = np.random.default_rng(42)
    n, g = 100_000, 6
    cond, batch = rng.integers(0, 2, n), rng.integers(0, 8, n)
    xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 8)[None, :])], axis=1).astype(float)
    zd = np.delete(xd, 1, axis=1)
    log_mu = np.log([0.02, 0.1, 0.3, 0.6, 1.0, 3.0])
    beta_b = rng.normal(0, 0.2, (8, g)); beta_b[0] = 0
    r_b = np.exp(rng.normal(0.5, 0.3, (8, g)))
    mu = np.exp(log_mu[None, :] + 0.4 * cond[:, None] * (np.arange(g) % 2)[None, :] + beta_b[batch])
    y = rng.poisson(rng.gamma(r_b[batch], mu / r_b[batch])).astype(np.float64)
    est = _fit(de, y, xd, zd)
    assert est._plan["kind"] == "grouped" and est.b_var.shape[0] == 8
    ref = onb.fit_nb(y, xd, zd, method_b="newton")
    sd = np.sqrt(np.einsum("gii->gi", ref["fisher_inv"])).T
    assert np.all(np.abs(est.a_var - ref["a_var"]) <= 1e-5 * np.abs(ref["a_var"]) + 2e-4 * sd)
    np.testing.assert_allclose(est.log_likelihood, ref["log_likelihood"], rtol=1e-9)
    np.testing.assert_array_equal(est.niter, ref["niter"])
    np.testing.assert_array_equal(est.error_codes, ref["error_codes"])
    # the dispersion of a batch is a flat direction of the likelihood where the batch holds a handful of counts (gene 0:
    # mean 0.02) or looks Poisson (r -> infinity): compare the per-batch log r where it is identified, as the model sees it
    zu = np.unique(zd, axis=0)
    logr, logr_ref = zu @ est.b_var, zu @ ref["b_var"]
    identified = (np.abs(logr_ref) < 8.0) & (np.arange(g)[None, :] >= 2)
    assert identified.mean() > 0.5
    np.testing.assert_allclose(logr[identified], logr_ref[identified], rtol=2e-3, atol=2e-3)


def test_niter_against_the_reference_faithful_brent_oracle(de, golden_dir):
    """How often the CUDA schedule (Newton scale step) takes exactly as many steps as the reference-faithful oracle
    (scipy Brent scale step) on the two frozen problems: reported, and required for the bulk of the genes.  The
    coefficients agree to the north-star tolerance on every gene off the Poisson plateau regardless."""
    report = {}
    for name in ("nbglm_cfg1.npz", "nbglm_batch.npz"):
        gd = np.load(os.path.join(golden_dir, name))
        est = _fit(de, gd["x"].astype(np.float64), gd["design_loc"], gd["design_scale"])
        frozen = (gd["newton_error_codes"] & onb.FLAG_SCALE_FROZEN) != 0
        same = est.niter == gd["brent_niter"]
        report[name] = (int(same[~frozen].sum()), int((~frozen).sum()), int(same[frozen].sum()), int(frozen.sum()))
        assert same[~frozen].mean() >= 0.9, "niter differs from the Brent oracle on more than 10 %% of the genes: %s" % (report,)
    print("identical niter vs Brent oracle (off plateau: same / genes, on plateau: same / genes):", report)
