"""
B-spline bases for ``de.test.continuous_1d`` without patsy: the construction of patsy's stateful transform
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
