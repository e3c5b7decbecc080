"""GPU (-m gpu): the gene-tiled per-cell K_fit (csrc/dx_fit_tile.cu) against the oracle, one test per kernel instantiation.

The launcher picks the instantiation from the number of entries of X^T W X (packed) + X^T S:
  p <= 7 -> <NE=1>, p <= 9 -> <2>, p <= 12 -> <3>, p <= 16 -> <5>, p <= 24 -> <11>, p <= 32 -> <18>,
each once for the constant dispersion model (ps = 1, the configs[3] case) and once for a general scale model.
Tolerances as in test_gpu_parity.py (north star): 1e-5 relative on coefficients, niter / flags bit-exact.
"""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import oracle.nb_glm as onb          # noqa: E402

from test_gpu_parity import _check_fit, _fit          # noqa: E402


@pytest.fixture(scope="module")
def de():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffxpy_b200.api as de_
    return de_


def _spline_like_design(rng, n, p):
    """Intercept + smooth bumps of a continuous covariate + a few batch indicators: full rank, entries in [0, 1]."""
    t = rng.uniform(0, 1, n)
    n_batch = min(3, max(p // 6, 0))
    n_bump = p - 1 - n_batch
    cols = [np.ones(n)]
    centers = np.linspace(0.0, 1.0, n_bump + 2)[1:-1] if n_bump > 0 else []
    width = 1.2 / max(n_bump, 1)
    for c in centers:
        cols.append(np.exp(-0.5 * ((t - c) / width) ** 2))
    if n_batch:
        batch = rng.integers(0, n_batch + 1, n)
        for b in range(1, n_batch + 1):
            cols.append((batch == b).astype(float))
    return np.stack(cols, axis=1), t


def _counts(rng, xd, sf, g, r_of_cell):
    p = xd.shape[1]
    beta = rng.normal(0, 0.25, (p, g))
    beta[0] = rng.uniform(np.log(0.5), np.log(15), g)
    mu = np.exp(xd @ beta) * sf[:, None]
    return rng.poisson(rng.gamma(r_of_cell, mu / r_of_cell)).astype(np.float64)


@pytest.mark.parametrize("p,n,g", [(3, 900, 37), (8, 1500, 33), (12, 2500, 20), (16, 3000, 9), (24, 4000, 6), (30, 5000, 4)])
def test_tile_kernel_constant_dispersion(de, p, n, g):
    """ps = 1 (constant dispersion) with per-cell size factors: every instantiation, a ragged last gene tile and cell tile."""
    rng = np.random.default_rng(100 + p)
    xd, _ = _spline_like_design(rng, n, p)
    sf = rng.uniform(0.5, 1.5, n)
    y = _counts(rng, xd, sf, g, 2.5)
    y[:, 0] = 0                              # an all-zero gene inside the tile
    est = _fit(de, y, xd, np.ones((n, 1)), size_factors=sf)
    assert est._plan["kind"] == "percell" and est.a_var.shape[0] == p
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), size_factors=sf, method_b="newton")
    skip = np.arange(g) == 0
    _check_fit(est, ref, flags_exact=True, skip=skip)
    np.testing.assert_allclose(est.b_var[:, ~skip], ref["b_var"][:, ~skip], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(est.jacobian[~skip], ref["jacobian"][~skip], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("p,ps,n,g", [(4, 2, 1200, 12), (9, 3, 2000, 10), (12, 8, 4000, 6), (16, 2, 3000, 5), (20, 2, 4000, 4),
                                       (28, 2, 5000, 3)])
def test_tile_kernel_general_scale_model(de, p, ps, n, g):
    """Scale model with numeric covariates (ps up to 8): the dispersion step goes through the contraction as well."""
    rng = np.random.default_rng(200 + p)
    xd, t = _spline_like_design(rng, n, p)
    zcols = [np.ones(n), t] + [np.cos((k + 1) * np.pi * t) * 0.5 for k in range(ps - 2)]
    zd = np.stack(zcols[:ps], axis=1)
    bz = np.concatenate([[0.6, 0.5], rng.normal(0, 0.15, max(ps - 2, 0))])[:ps]
    r = np.exp(zd @ bz)[:, None]
    sf = np.ones(n)
    y = _counts(rng, xd, sf, g, r)
    est = _fit(de, y, xd, zd)
    assert est._plan["kind"] == "percell" and est.b_var.shape[0] == ps
    ref = onb.fit_nb(y, xd, zd, method_b="newton")
    _check_fit(est, ref, flags_exact=True)
    np.testing.assert_allclose(est.b_var, ref["b_var"], rtol=2e-5, atol=5e-6)


def test_tile_kernel_high_counts_and_noninteger_values(de):
    """Counts beyond the per-gene tables (y >= 64) and non-integer values take the per-cell lgamma / psi path."""
    rng = np.random.default_rng(7)
    n, g, p = 2000, 8, 5
    xd, _ = _spline_like_design(rng, n, p)
    sf = rng.uniform(0.7, 1.3, n)
    beta = rng.normal(0, 0.2, (p, g))
    beta[0] = rng.uniform(np.log(20), np.log(300), g)
    mu = np.exp(xd @ beta) * sf[:, None]
    y = rng.poisson(rng.gamma(3.0, mu / 3.0)).astype(np.float64)
    y[:, 1] = np.round(y[:, 1] * 0.5, 1)          # non-integer "normalised" values
    est = _fit(de, y, xd, np.ones((n, 1)), size_factors=sf)
    ref = onb.fit_nb(y.astype(np.float32).astype(np.float64), xd, np.ones((n, 1)), size_factors=sf, method_b="newton")
    _check_fit(est, ref, flags_exact=True)


def test_tile_kernel_full_cell_count_configs3_shape(de):
    """configs[3] shape at the full cell count: 250k cells, 12 spline-like coefficients, size factors, a few genes."""
    rng = np.random.default_rng(33)
    n, g, p = 250_000, 4, 12
    xd, _ = _spline_like_design(rng, n, p)
    sf = rng.uniform(0.5, 1.5, n)
    beta = rng.normal(0, 0.3, (p, g))
    beta[0] = np.log([0.05, 0.3, 1.0, 4.0])
    mu = np.exp(xd @ beta) * sf[:, None]
    y = rng.poisson(rng.gamma(1.5, mu / 1.5)).astype(np.float64)
    est = _fit(de, y, xd, np.ones((n, 1)), size_factors=sf)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), size_factors=sf, method_b="newton")
    # full size: same allowance as test_cfg2_full_cell_count_against_oracle (the last step of a gene is accepted or
    # reverted on a likelihood difference at rounding level)
    sd = np.sqrt(np.einsum("gii->gi", ref["fisher_inv"])).T
    assert np.all(np.abs(est.a_var - ref["a_var"]) <= 1e-5 * np.abs(ref["a_var"]) + 2e-4 * sd)
    np.testing.assert_allclose(est.log_likelihood, ref["log_likelihood"], rtol=1e-9)
    np.testing.assert_array_equal(est.niter, ref["niter"])
    np.testing.assert_array_equal(est.error_codes, ref["error_codes"])


def test_tile_kernel_rejects_more_than_8_scale_coefficients(de):
    from diffxpy_b200 import _lib
    rng = np.random.default_rng(1)
    n, g = 400, 2
    y = rng.poisson(3.0, (n, g)).astype(np.float64)
    xd = np.stack([np.ones(n), rng.uniform(0, 1, n)], axis=1)
    zd = np.concatenate([np.ones((n, 1)), rng.uniform(0, 1, (n, 9))], axis=1)
    with pytest.raises((_lib.DxError, ValueError)):
        _fit(de, y, xd, zd)
