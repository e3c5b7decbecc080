"""
Spline bases for ``de.test.continuous_1d`` without patsy (B-splines here, cubic regression splines further down): the construction of patsy's stateful transform
``bs(x, df=df, degree=3, include_intercept=False)`` that the reference requests at
/root/reference/diffxpy/testing/tests.py:2196-2200 (inner knots at equally spaced quantiles of x, boundary knots at
min / max repeated degree + 1 times, first basis column dropped), evaluated with scipy.
"""
import numpy as np
from scipy.interpolate import BSpline


def bs_knots(x, df, degree=3, include_intercept=False):
    """All knots (boundary knots repeated ``degree + 1`` times) of patsy's ``bs`` for the sample ``x``."""
    x = np.asarray(x, dtype=np.float64)
    order = degree + 1
    n_inner = df - order + (0 if include_intercept else 1)
    if n_inner < 0:
        raise ValueError("df=%r is too small for degree=%r; must be >= %s" % (df, degree, order - (0 if include_intercept else 1)))
    probs = np.linspace(0, 1, n_inner + 2)[1:-1]
    inner = np.quantile(x, probs) if n_inner > 0 else np.zeros(0)        # linear interpolation (R type 7), like patsy
    lo, hi = float(np.min(x)), float(np.max(x))
    knots = np.concatenate([[lo, hi] * order, inner])
    knots.sort()
    return knots


def bs_basis(x, knots, degree=3, include_intercept=False):
    """Basis matrix (len(x) x df) on the given knots; values outside the boundary knots are an error (as in patsy)."""
    x = np.asarray(x, dtype=np.float64)
    lo, hi = knots[0], knots[-1]
    if np.any(x < lo) or np.any(x > hi):
        raise NotImplementedError("some data points fall outside the outermost knots, and extrapolation is not supported")
    basis = BSpline.design_matrix(x, knots, degree).toarray()
    return basis if include_intercept else basis[:, 1:]


def bs(x, df, degree=3, include_intercept=False):
    """patsy ``bs(x, df=df, degree=degree, include_intercept=include_intercept)`` on its own sample."""
    return bs_basis(x, bs_knots(x, df, degree, include_intercept), degree, include_intercept)


# ---------------------------------------------------------------------------------------------------------
# Natural ("cr") and cyclic ("cc") cubic regression splines with the centering constraint, i.e. patsy's
# ``cr(x, df=df, constraints='center')`` / ``cc(x, df=df, constraints='center')`` that the reference requests at
# /root/reference/diffxpy/testing/tests.py:2201-2210.  patsy follows mgcv (Wood, Generalized Additive Models, 4.1.2):
# a cardinal basis (one function per knot, equal to 1 at its knot and 0 at the others) of the cubic splines on the
# knots that are linear beyond the boundary knots (cr) or periodic over them (cc); the constraint "every column has
# mean zero over the sample" is absorbed with the QR factorisation of the constraint row, which removes one column.
# Knots: df + 1 (cr) or df + 2 (cc) of them, the inner ones at equally spaced quantiles of the UNIQUE sample values,
# the outer ones at min / max.  patsy itself is not importable here, so the construction is pinned by its defining
# properties in tests/test_host_logic.py (interpolation at the knots, C2 continuity, natural / periodic boundary
# conditions, zero column means, dimension) -- they fix the spanned function space, hence every test statistic of a
# model that contains an intercept; single coefficients are only comparable if the basis matches column by column.
# ---------------------------------------------------------------------------------------------------------
def crs_knots(x, df, cyclic=False, n_constraints=1):
    """All knots of patsy's cr / cc transform for the sample ``x`` (constraints='center' -> n_constraints = 1)."""
    x = np.asarray(x, dtype=np.float64)
    n_inner = df - (1 if cyclic else 2) + n_constraints
    if n_inner < 0:
        raise ValueError("df=%r is too small for the spline basis" % df)
    ux = np.unique(x)
    probs = np.linspace(0, 100, n_inner + 2)[1:-1]
    inner = np.percentile(ux, probs.tolist()) if n_inner > 0 else np.zeros(0)
    knots = np.unique(np.concatenate([[ux[0], ux[-1]], inner]))
    if knots.size != n_inner + 2:
        raise ValueError("the sample has too few distinct values for df=%r" % df)
    return knots


def _crs_second_derivative_map(knots, cyclic):
    """Matrix F with (second derivatives at the knots) = F @ (values at the knots) for the interpolating spline."""
    h = np.diff(knots)
    n = knots.size
    if not cyclic:
        # natural spline: zero second derivative at both boundary knots; tridiagonal system for the inner ones
        b = np.zeros((n - 2, n - 2))
        d = np.zeros((n - 2, n))
        for i in range(n - 2):
            b[i, i] = (h[i] + h[i + 1]) / 3.0
            if i + 1 < n - 2:
                b[i, i + 1] = b[i + 1, i] = h[i + 1] / 6.0
            d[i, i] = 1.0 / h[i]
            d[i, i + 2] = 1.0 / h[i + 1]
            d[i, i + 1] = -d[i, i] - d[i, i + 2]
        f = np.zeros((n, n))
        f[1:-1] = np.linalg.solve(b, d)
        return f
    # periodic spline on n - 1 distinct knots (the last knot is the first one again)
    m = n - 1
    b = np.zeros((m, m))
    d = np.zeros((m, m))
    for i in range(m):
        hm, hp = h[i - 1], h[i]                 # interval before / after knot i (h[-1] wraps around)
        b[i, i] += (hm + hp) / 3.0
        b[i, (i - 1) % m] += hm / 6.0
        b[i, (i + 1) % m] += hp / 6.0
        d[i, i] += -1.0 / hm - 1.0 / hp
        d[i, (i - 1) % m] += 1.0 / hm
        d[i, (i + 1) % m] += 1.0 / hp
    return np.linalg.solve(b, d)


def crs_free_basis(x, knots, cyclic=False):
    """Unconstrained cardinal basis (len(x) x n_knots, or n_knots - 1 columns for the cyclic one)."""
    x = np.asarray(x, dtype=np.float64)
    knots = np.asarray(knots, dtype=np.float64)
    n = knots.size
    lo, hi = knots[0], knots[-1]
    if cyclic:
        x = lo + np.mod(x - lo, hi - lo)
    j = np.clip(np.searchsorted(knots, x) - 1, 0, n - 2)
    h = np.diff(knots)
    hj = h[j]
    up, dn = knots[j + 1] - x, x - knots[j]
    ajm, ajp = up / hj, dn / hj
    cjm3 = up ** 3 / (6.0 * hj)
    cjp3 = dn ** 3 / (6.0 * hj)
    cjm3[x > hi] = 0.0                          # linear continuation beyond the boundary knots
    cjp3[x < lo] = 0.0
    cjm = cjm3 - hj * up / 6.0
    cjp = cjp3 - hj * dn / 6.0
    f = _crs_second_derivative_map(knots, cyclic)
    ncol = n - 1 if cyclic else n
    j1 = j + 1
    if cyclic:
        j1 = np.where(j1 == n - 1, 0, j1)
    eye = np.identity(ncol)
    return ajm[:, None] * eye[j] + ajp[:, None] * eye[j1] + cjm[:, None] * f[j] + cjp[:, None] * f[j1]


def crs(x, df, cyclic=False, center=True, knots=None, constraint=None):
    """patsy ``cr`` / ``cc`` (``cyclic``) with ``df`` columns on its own sample; ``center`` = constraints='center'.
    Returns (basis, knots, constraint row) so that the same transform can be applied to new points."""
    x = np.asarray(x, dtype=np.float64)
    if knots is None:
        knots = crs_knots(x, df, cyclic=cyclic, n_constraints=1 if center else 0)
    free = crs_free_basis(x, knots, cyclic=cyclic)
    if not center:
        return free, knots, None
    if constraint is None:
        constraint = np.atleast_2d(free.mean(axis=0))
    q, _ = np.linalg.qr(constraint.T, mode="complete")
    return free @ q[:, constraint.shape[0]:], knots, constraint
