"""GPU (-m gpu): the round-2 kernels, called through the C ABI.

* dx_fit_lane_kernel (lane-resident K_fit, dx_fit_lane.cu) against the general K_fit kernel on the same
  sufficient statistics (niter / flags identical, coefficients to rounding) and against the CPU oracle
  (oracle/nb_glm.py, method_b="newton": north-star tolerances), for every instantiation: 1 / 2 / 4 / 8 / 16 lanes per
  gene (DXB200_FIT_LPG), coefficient blocks 4 / 9, with size-factor groups, overflow lists (plain and packed) and empty genes;
* the Wald row the fit kernels write in their epilogue against dx_wald_from_fit and the reference golden rule for a
  singular Fisher information (NaN, not p = 0);
* the single-CTA Benjamini-Hochberg kernel against the oracle BH (bit for bit) and the multi-kernel path.
"""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import oracle.nb_glm as onb          # noqa: E402
import oracle.stats as ost           # noqa: E402


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffxpy_b200 import _lib
    _lib.require_cuda()
    return _lib


def _simulate(rng, n, g, cond, batch, n_batch, mu_lo, mu_hi, sf=None):
    log_mu = rng.uniform(np.log(mu_lo), np.log(mu_hi), g)
    r = rng.uniform(0.5, 4.0, g)
    beta_c = np.where(rng.uniform(size=g) < 0.3, rng.normal(0, 0.5, g), 0.0)
    beta_b = rng.normal(0, 0.2, (n_batch, g))
    beta_b[0] = 0.0
    eta = log_mu[None, :] + cond[:, None] * beta_c[None, :] + beta_b[batch]
    if sf is not None:
        eta = eta + np.log(sf)[:, None]
    return rng.poisson(rng.gamma(r[None, :], np.exp(eta) / r[None, :])).astype(np.float64)


def _design(cond, batch, n_batch):
    n = cond.shape[0]
    return np.concatenate([np.ones((n, 1)), cond[:, None].astype(np.float64),
                           (batch[:, None] == np.arange(1, n_batch)[None, :]).astype(np.float64)], axis=1)


def _run_fit(lib, y, xd, log_sf=None, scale_const=True, pack=True, coef=1, init_a=None, train_loc=1):
    """ingest + (pack) + dx_nb_fit_wald_grouped on one GPU; returns a dict of host arrays."""
    from diffxpy_b200.device import CountShard, ptr, stream_ptr
    from diffxpy_b200.estimator import group_rows
    dev = torch.device("cuda")
    n, g = y.shape
    p = xd.shape[1]
    cols = [xd, np.ones((n, 1))]
    if log_sf is not None:
        cols.append(log_sf.reshape(-1, 1))
    uniq, inv = group_rows(np.concatenate(cols, axis=1))
    k = uniq.shape[0]
    assert k <= 255
    xu = np.ascontiguousarray(uniq[:, :p])
    xs = np.ascontiguousarray(uniq[:, p:p + 1])
    off = np.ascontiguousarray(uniq[:, p + 1]) if log_sf is not None else None
    nk = np.bincount(inv, minlength=k).astype(np.float64)
    loc_u, loc_inv = group_rows(xd)
    first = np.zeros(k, dtype=np.int64)
    first[inv[::-1]] = np.arange(n - 1, -1, -1)
    loc_group = loc_inv[first].astype(np.int32)
    pinv = np.ascontiguousarray(np.linalg.pinv(loc_u))
    shard = CountShard.from_host(scipy.sparse.csc_matrix(y.astype(np.float32)), dev)
    suff = shard.group_suffstats(torch.from_numpy(inv.astype(np.uint8)).to(dev), k)
    if pack:
        suff.pack_overflow()
    one_hot = set(np.unique(xd)) <= {0.0, 1.0}
    if init_a is None:
        init_a = lib.DX_INIT_CLOSED_FORM if one_hot else lib.DX_INIT_STANDARD
    opts = lib.DxFitOpts(max_steps=1000, update_b_freq=5, train_loc=int(train_loc), train_scale=1, init_a=init_a,
                         init_b=lib.DX_INIT_STANDARD, newton_max_iter=100,
                         reserved=lib.DX_OPT_SCALE_CONST if scale_const else 0, lltol=1e-10, newton_xtol=1e-8)
    f64 = dict(dtype=torch.float64, device=dev)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = dict(xu=t(xu), xs=t(xs), off=t(off), nk=t(nk), pinv=t(pinv), lg=t(loc_group))
    out = dict(a=torch.zeros((p, g), **f64), b=torch.zeros((1, g), **f64), fi=torch.zeros((g, p, p), **f64),
               ll=torch.zeros(g, **f64), jac=torch.zeros(g, **f64), ni=torch.zeros(g, dtype=torch.int32, device=dev),
               fl=torch.zeros(g, dtype=torch.int32, device=dev), th=torch.zeros(g, **f64), sd=torch.zeros(g, **f64),
               pv=torch.zeros(g, **f64), lp=torch.zeros(g, **f64))
    lib.call("dx_nb_fit_wald_grouped", g, k, p, 1, ptr(d["xu"]), ptr(d["xs"]), ptr(d["off"]), ptr(d["nk"]), ptr(d["pinv"]),
             ptr(d["lg"]), int(loc_u.shape[0]), ptr(suff.tail), ptr(suff.ovf_count), ptr(shard.indptr), ptr(suff.ovf_val),
             ptr(suff.ovf_group), ptr(suff.ovf_packed) if pack else None, opts,
             ptr(out["a"]), ptr(out["b"]), ptr(out["fi"]), ptr(out["ll"]), ptr(out["jac"]), ptr(out["ni"]), ptr(out["fl"]),
             coef, ptr(out["th"]), ptr(out["sd"]), ptr(out["pv"]), ptr(out["lp"]), None, 1, 0, 0, 0, stream_ptr())
    torch.cuda.synchronize()
    res = {key: v.cpu().numpy() for key, v in out.items()}
    res["k"] = k
    return res


REGULAR_MAX_STEPS = 60


def _assert_same_fit(fast, gen):
    """register-resident kernel vs general kernel on identical statistics: same decisions, values to the accuracy the
    stopping rule leaves (relative log-likelihood change 1e-10 -> coefficients to ~1e-6).
    Genes outside the regular regime are compared on their likelihood only: a design group without counts has no finite
    maximum (the coefficient drifts towards -inf for hundreds of steps and the stopping step, or whether the Fisher
    information counts as singular on the way, depends on rounding), and on the Poisson plateau (b beyond ~20) the
    likelihood is flat in the dispersion."""
    reg = (gen["ni"] <= REGULAR_MAX_STEPS) & (fast["ni"] <= REGULAR_MAX_STEPS) & (gen["fl"] == 1) & (np.abs(gen["b"][0]) < 20)
    assert reg.mean() > 0.8
    np.testing.assert_array_equal(fast["ni"][reg], gen["ni"][reg])
    np.testing.assert_array_equal(fast["fl"][reg], gen["fl"][reg])
    np.testing.assert_allclose(fast["ll"], gen["ll"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(fast["ll"][reg], gen["ll"][reg], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(fast["a"][:, reg], gen["a"][:, reg], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(fast["b"][:, reg], gen["b"][:, reg], rtol=1e-5, atol=1e-5)
    scale = np.abs(gen["fi"]).max(axis=(1, 2), keepdims=True)
    both_nan = np.isnan(fast["fi"]) & np.isnan(gen["fi"])          # (a singular Fisher information is NaN in both)
    scale = np.nan_to_num(scale, nan=1.0)
    assert np.all((both_nan | (np.abs(fast["fi"] - gen["fi"]) <= 1e-4 * scale + 1e-300))[reg])      # (the dispersion agrees to 1e-5 + 1e-5 b)
    np.testing.assert_allclose(fast["sd"][reg], gen["sd"][reg], rtol=1e-6)
    ok = reg & (gen["pv"] > 1e-11)
    np.testing.assert_allclose(fast["lp"][ok], gen["lp"][ok], rtol=1e-5, atol=1e-9)


def _assert_oracle(res, ref, coef=1):
    # regular genes: converged in a few steps and off the Poisson plateau (beyond b ~ 10 the likelihood is flat in the
    # dispersion to 1e-9 relative and the stopping step of either implementation is decided by rounding)
    reg = (ref["niter"] <= REGULAR_MAX_STEPS) & (res["ni"] <= REGULAR_MAX_STEPS) & (ref["error_codes"] == 1) & \
        (np.abs(ref["b_var"][0]) < 10) & (np.abs(res["b"][0]) < 10) & (np.abs(ref["a_var"]).max(axis=0) < 50)
    # (last condition: a design group without counts sends coefficients to the parameter bounds, +-290; what the schedule
    # reports for such a gene is decided by rounding)
    assert reg.mean() > 0.8
    np.testing.assert_array_equal(res["ni"][reg], ref["niter"][reg])
    np.testing.assert_array_equal(res["fl"][reg], ref["error_codes"][reg])
    np.testing.assert_allclose(res["a"][:, reg], ref["a_var"][:, reg], rtol=1e-5, atol=2e-6)
    # (1e-8: the last scale step of a gene improves the likelihood by less than its rounding noise, and the "revert if
    # worse" rule may resolve differently for the two summation orders -- DESIGN.md section 2)
    np.testing.assert_allclose(res["ll"][reg], ref["log_likelihood"][reg], rtol=1e-8, atol=1e-9)
    # outside the regular regime (a design group without counts: the coefficient drifts towards -inf and the step at which
    # the schedule stops depends on rounding) the two paths stop at slightly different points of the same flat ridge
    sane = np.abs(ref["a_var"]).max(axis=0) < 50
    np.testing.assert_allclose(res["ll"][sane], ref["log_likelihood"][sane], rtol=1e-4, atol=1e-3)
    fi_ref = ref["fisher_inv"]
    scale = np.abs(fi_ref).max(axis=(1, 2), keepdims=True)
    assert np.all((np.abs(res["fi"] - fi_ref) <= 2e-5 * scale + 1e-300)[reg])
    sd_ref = np.sqrt(np.maximum(fi_ref[:, coef, coef], np.nextafter(0, 1)))
    with np.errstate(all="ignore"):
        p_ref = ost.wald_test(ref["a_var"][coef], sd_ref)
    ok = reg & (p_ref > 1e-11)
    # (atol: a coefficient within its 2e-6 tolerance of zero has log p ~ -0.8 |z|, relative to which any error is large)
    np.testing.assert_allclose(np.log(res["pv"][ok]), np.log(p_ref[ok]), rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("n_batch,lpg", [(8, 0), (8, 1), (8, 2), (8, 4), (8, 8), (8, 16), (3, 1), (3, 2), (3, 4), (3, 16),
                                         (4, 1), (4, 8), (1, 1), (1, 2), (12, 0)])
def test_fast_kernel_vs_general_kernel_and_oracle(lib, n_batch, lpg, monkeypatch):
    """'~ 1 + condition + batch' with 2 x n_batch design groups: K = 16 / p = 9 is the headline case (coefficient block 9)
    with every lanes-per-gene instantiation (0 = the launcher's own choice); n_batch = 3 -> K = 6, p = 4 (block 4);
    n_batch = 4 -> K = 8, p = 5 (block 9, partly filled); n_batch = 1 -> two groups, p = 2 (BASELINE configs[0]: the closed-form
    start is the maximum, so the reference does not train the location model);
    n_batch = 12 -> K = 24, p = 13 is beyond the lane-resident kernel: the general one (16 lanes striding over 24 groups) runs
    both times and is checked against the oracle."""
    if lpg:
        monkeypatch.setenv("DXB200_FIT_LPG", str(lpg))
    rng = np.random.default_rng(100 + n_batch)
    n, g = 3000, 96
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, n_batch, n)
    y = _simulate(rng, n, g, cond, batch, n_batch, 0.02, 12.0)
    y[:, 5] = 0.0                                        # all-zero gene
    y[:, 6] = 0.0
    y[7, 6] = 3.0                                        # single non-zero
    xd = _design(cond, batch, n_batch)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
    fast = _run_fit(lib, y, xd, scale_const=True, train_loc=ref["train_loc"])
    gen = _run_fit(lib, y, xd, scale_const=False, train_loc=ref["train_loc"])
    assert fast["k"] == 2 * n_batch
    _assert_oracle(fast, ref)
    _assert_oracle(gen, ref)
    _assert_same_fit(fast, gen)
    # the epilogue row == dx_wald_from_fit on the fit's own outputs
    from diffxpy_b200.device import ptr, stream_ptr
    dev = torch.device("cuda")
    a_d, fi_d = torch.from_numpy(fast["a"]).to(dev), torch.from_numpy(fast["fi"]).to(dev)
    rows = torch.empty((4, g), dtype=torch.float64, device=dev)
    lib.call("dx_wald_from_fit", g, xd.shape[1], 1, ptr(a_d), ptr(fi_d), ptr(rows[0]), ptr(rows[1]), ptr(rows[2]), ptr(rows[3]),
             stream_ptr())
    r = rows.cpu().numpy()
    np.testing.assert_array_equal(r[0], fast["th"])
    np.testing.assert_array_equal(r[1], fast["sd"])
    np.testing.assert_array_equal(r[2], fast["pv"])
    np.testing.assert_array_equal(r[3], fast["lp"])


@pytest.mark.parametrize("lpg", [0, 1, 4, 16])
def test_fast_kernel_size_factor_groups_and_overflow(lib, lpg, monkeypatch):
    """Per-group size factors (offset in the linear predictor, 3 levels x 2 conditions x 2 batches = 12 groups) and highly
    expressed genes: counts above 64 go through the overflow lists, plain and packed (K_pack records)."""
    if lpg:
        monkeypatch.setenv("DXB200_FIT_LPG", str(lpg))
    rng = np.random.default_rng(7)
    n, g = 2400, 64
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 2, n)
    sf = rng.choice([0.5, 1.0, 2.0], size=n)
    y = _simulate(rng, n, g, cond, batch, 2, 0.5, 400.0, sf=sf)
    xd = _design(cond, batch, 2)
    for pack in (True, False):
        fast = _run_fit(lib, y, xd, log_sf=np.log(sf), scale_const=True, pack=pack, init_a=lib.DX_INIT_STANDARD)
        gen = _run_fit(lib, y, xd, log_sf=np.log(sf), scale_const=False, pack=pack, init_a=lib.DX_INIT_STANDARD)
        assert fast["k"] == 12
        _assert_same_fit(fast, gen)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), size_factors=sf, method_b="newton", init_a="standard")
    _assert_oracle(fast, ref)


@pytest.mark.parametrize("lpg", [0, 1, 8, 16])
def test_lane_kernel_32_groups(lib, lpg, monkeypatch):
    """2 conditions x 8 batches x 2 size-factor levels = 32 design groups at 9 coefficients: two to 32 groups per lane."""
    if lpg:
        monkeypatch.setenv("DXB200_FIT_LPG", str(lpg))
    rng = np.random.default_rng(13)
    n, g = 6000, 80
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 8, n)
    sf = rng.choice([0.7, 1.4], size=n)
    y = _simulate(rng, n, g, cond, batch, 8, 0.05, 20.0, sf=sf)
    xd = _design(cond, batch, 8)
    fast = _run_fit(lib, y, xd, log_sf=np.log(sf), scale_const=True)
    gen = _run_fit(lib, y, xd, log_sf=np.log(sf), scale_const=False)
    assert fast["k"] == 32
    _assert_same_fit(fast, gen)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), size_factors=sf, method_b="newton")
    _assert_oracle(fast, ref)


@pytest.mark.parametrize("lpg", [0, 1, 2])
def test_fast_kernel_numeric_design_dense_products(lib, lpg, monkeypatch):
    """A design whose rows are not 0/1 (numeric covariate with 5 levels x condition): no zero entries to skip; standard
    init."""
    if lpg:
        monkeypatch.setenv("DXB200_FIT_LPG", str(lpg))
    rng = np.random.default_rng(11)
    n, g = 2000, 48
    cond = rng.integers(0, 2, n)
    dose = rng.choice([0.0, 0.5, 1.0, 1.5, 2.5], size=n)
    xd = np.stack([np.ones(n), cond.astype(np.float64), dose, dose * cond], axis=1)
    log_mu = rng.uniform(np.log(0.1), np.log(8.0), g)
    eta = log_mu[None, :] + 0.3 * cond[:, None] + 0.2 * dose[:, None]
    r = rng.uniform(0.5, 3.0, g)
    y = rng.poisson(rng.gamma(r[None, :], np.exp(eta) / r[None, :])).astype(np.float64)
    fast = _run_fit(lib, y, xd, scale_const=True, coef=2)
    gen = _run_fit(lib, y, xd, scale_const=False, coef=2)
    assert fast["k"] == 10
    _assert_same_fit(fast, gen)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
    _assert_oracle(fast, ref, coef=2)


def test_singular_fisher_information_gives_nan_row_not_p_zero(lib):
    """A design group without cells of one condition makes X^T W X singular for that ... here: a duplicated design column.
    The final factorisation fails, fisher_inv is NaN, and the Wald row must be NaN (the reference's nextafter floor leaves
    NaN alone, stats.py:215) and be excluded from the q-values."""
    from diffxpy_b200.testing import correction
    rng = np.random.default_rng(3)
    n, g = 1500, 40
    cond = rng.integers(0, 2, n)
    xd = np.stack([np.ones(n), cond.astype(np.float64), cond.astype(np.float64)], axis=1)      # rank deficient on purpose
    y = rng.poisson(1.5, size=(n, g)).astype(np.float64)
    for sc in (True, False):
        res = _run_fit(lib, y, xd, scale_const=sc, init_a=lib.DX_INIT_STANDARD)
        assert np.all(res["fl"] & lib.FLAG_SINGULAR_FIM)
        assert np.all(np.isnan(res["fi"])) and np.all(np.isnan(res["pv"])) and np.all(np.isnan(res["lp"]))
        assert np.all(np.isnan(res["sd"]))
        q = correction.correct(np.concatenate([res["pv"], [0.01, 0.5]]))
        assert np.all(np.isnan(q[:g])) and np.all(np.isfinite(q[g:]))


def test_fused_bh_kernel_bitwise(lib):
    """n <= 32768 runs in one CTA: identical q-values to the oracle BH, the multi-kernel path and scipy."""
    from diffxpy_b200.testing import correction
    rng = np.random.default_rng(21)
    for n in (1, 2, 31, 1024, 1025, 20000, 32768):
        p = rng.uniform(size=n) ** 3
        if n > 100:
            p[::53] = np.nan
            p[1::71] = 0.0
            p[2::67] = 1.0
            p[3::11] = np.round(p[3::11], 2)             # ties
            p[5::101] = 10.0 ** rng.uniform(-300, -5, p[5::101].shape[0])
        want = ost.bh_correct(p)
        got = correction.correct(p)
        np.testing.assert_array_equal(got, want)
    # the multi-kernel path takes over above 32768 p-values
    p = rng.uniform(size=40000)
    p[::97] = np.nan
    np.testing.assert_array_equal(correction.correct(p), ost.bh_correct(p))


def test_host_entry_point_returns_q_values(lib):
    """dx_wald_grouped_host: q-values of all genes come back with the rows (det.py:108-117 is part of the Wald result)."""
    from diffxpy_b200.estimator import group_rows
    rng = np.random.default_rng(5)
    n, g = 2500, 300
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 4, n)
    y = _simulate(rng, n, g, cond, batch, 4, 0.05, 5.0)
    xd = _design(cond, batch, 4)
    p = xd.shape[1]
    uniq, inv = group_rows(np.concatenate([xd, np.ones((n, 1))], axis=1))
    k = uniq.shape[0]
    csc = scipy.sparse.csc_matrix(y.astype(np.float32))
    indptr = np.ascontiguousarray(csc.indptr, dtype=np.int64)
    rows = np.ascontiguousarray(csc.indices, dtype=np.int32)
    vals = np.ascontiguousarray(csc.data, dtype=np.float32)
    grp = inv.astype(np.uint8)
    xu = np.ascontiguousarray(uniq[:, :p]); xs = np.ascontiguousarray(uniq[:, p:p + 1])
    nk = np.bincount(inv, minlength=k).astype(np.float64)
    pinv = np.ascontiguousarray(np.linalg.pinv(xu))
    opts = lib.DxFitOpts(max_steps=1000, update_b_freq=5, train_loc=1, train_scale=1, init_a=lib.DX_INIT_CLOSED_FORM,
                         init_b=lib.DX_INIT_STANDARD, newton_max_iter=100, reserved=lib.DX_OPT_SCALE_CONST, lltol=1e-10,
                         newton_xtol=1e-8)
    a = np.zeros((p, g)); b = np.zeros((1, g)); fi = np.zeros((g, p, p)); ll = np.zeros(g); jac = np.zeros(g)
    ni = np.zeros(g, dtype=np.int32); fl = np.zeros(g, dtype=np.int32)
    pv = np.zeros(g); lp = np.zeros(g); sdv = np.zeros(g); q = np.zeros(g)
    P = lambda arr: arr.ctypes.data_as(C.c_void_p)
    for chunks in ("1", "4"):
        os.environ["DXB200_HOST_CHUNKS"] = chunks
        try:
            lib.call("dx_wald_grouped_host", n, g, P(indptr), P(rows), P(vals), P(grp), k, p, 1, P(xu), P(xs), None, P(nk),
                     P(pinv), None, k, C.byref(opts), 1, P(a), P(b), P(fi), P(ll), P(jac), P(ni), P(fl), P(pv), P(lp), P(sdv),
                     P(q))
        finally:
            del os.environ["DXB200_HOST_CHUNKS"]
        np.testing.assert_array_equal(q, ost.bh_correct(pv))
        dev = _run_fit(lib, y, xd, scale_const=True)
        np.testing.assert_array_equal(a, dev["a"])
        np.testing.assert_array_equal(pv, dev["pv"])
        np.testing.assert_array_equal(sdv, dev["sd"])
    lib.call("dx_host_release")
    lib.call("dx_host_release")          # idempotent


@pytest.mark.parametrize("k", [2, 16, 32])
def test_ingest_split_mode_equals_gene_mode(lib, k, monkeypatch):
    """K_ingest cuts a shard with few genes into equal ranges of non-zeros (pieces of genes, tail counts added with
    atomics) and redoes the genes with overflow entries gene-wise: same tail counts, same overflow lists in the same
    (canonical) order as the one-warp-per-gene mode, and as numpy."""
    from diffxpy_b200.device import CountShard
    rng = np.random.default_rng(17 + k)
    n = 60000
    lens = [0, 1, 3, 255, 256, 257, 1023, 1024, 1025, 0, 0, 4099, 8191, 8192, 30000, 59999, 7, 0, 12345, 2, 40000, 0]
    cols, rws, vls = [], [], []
    for gi, m in enumerate(lens):
        r = np.sort(rng.choice(n, size=m, replace=False))
        v = rng.integers(1, 12, size=m).astype(np.float32)
        if gi % 3 == 0 and m > 100:
            v[::97] = 70.0          # above the direct bins
            v[5::89] = 2.5          # non-integer
        cols.append(np.full(m, gi)); rws.append(r); vls.append(v)
    x = scipy.sparse.csc_matrix((np.concatenate(vls), (np.concatenate(rws), np.concatenate(cols))), shape=(n, len(lens)))
    dense = np.asarray(x.todense())
    grp = rng.integers(0, k, n).astype(np.uint8)
    grp[::501] = 255
    grp_d = torch.from_numpy(grp).cuda()
    res = {}
    # "0" / "1": the gene-wise entry point in its one-warp-per-gene and equal-ranges modes; "p": the work-list entry point
    # (dx_group_suffstats_pieces, what CountShard.group_suffstats calls by default: here the genes above 6144 entries are cut)
    for mode in ("0", "1", "p"):
        if mode == "p":
            monkeypatch.delenv("DXB200_INGEST_SPLIT")
            monkeypatch.delenv("DXB200_INGEST_PIECES")
        else:
            monkeypatch.setenv("DXB200_INGEST_SPLIT", mode)
            monkeypatch.setenv("DXB200_INGEST_PIECES", "0")
        shard = CountShard.from_host(x)
        if mode == "p":
            pieces, multi = shard.work_list
            pc = pieces.cpu().numpy()
            assert multi.shape[0] >= 5 and pc[:, 3].sum() > multi.shape[0]          # several genes are cut into parts
            assert np.array_equal(np.bincount(pc[:, 0], weights=pc[:, 2], minlength=len(lens)).astype(np.int64), np.array(lens))
        suff = shard.group_suffstats(grp_d, k)
        torch.cuda.synchronize()
        res[mode] = (suff.tail.cpu().numpy().view(np.uint32), suff.ovf_count.cpu().numpy(), suff.ovf_val.cpu().numpy(),
                     suff.ovf_group.cpu().numpy(), shard.indptr.cpu().numpy())
    indptr = res["0"][4]
    for other in ("1", "p"):
        np.testing.assert_array_equal(res["0"][0], res[other][0])
        np.testing.assert_array_equal(res["0"][1], res[other][1])
        for gi in range(len(lens)):
            c = res["0"][1][gi]
            sl = slice(indptr[gi], indptr[gi] + c)
            np.testing.assert_array_equal(res["0"][2][sl], res[other][2][sl])
            np.testing.assert_array_equal(res["0"][3][sl], res[other][3][sl])
    for gi in range(len(lens)):
        c = res["0"][1][gi]
        col = dense[:, gi]
        isbin = (col >= 1) & (col <= 64) & (col == np.floor(col))
        for kk in range(0, k, max(1, k // 4)):
            sel = isbin & (grp == kk)
            want = np.array([(col[sel] > j).sum() for j in range(64)])
            np.testing.assert_array_equal(res["1"][0][gi, kk], want, err_msg="gene %d group %d" % (gi, kk))
        assert c == ((col != 0) & ~isbin & (grp != 255)).sum()
