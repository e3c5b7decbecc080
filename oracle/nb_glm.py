"""
CPU oracle for the per-gene negative-binomial GLM fit.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).  **PARITY UNPINNED** against real batchglm (not installable offline, and the reference's
tests hold no fitted number); the END POINT of the schedule is pinned independently: tests/test_oracle_golden.py
maximises the same likelihood with scipy BFGS and requires likelihood, coefficients, dispersion and Fisher inverse
to coincide.

What is restated (numpy, float64, dense cells x genes like the reference's numpy backend):

* model            eta_loc = design_loc @ (constraints_loc @ a) + log(size_factors), loc = exp(eta_loc)
                   eta_scale = design_scale @ (constraints_scale @ b), scale = exp(eta_scale)   (NB size)
                   -- parametrisation confirmed by /root/reference/diffxpy/testing/det.py:1262-1265
                   and det_cont.py:221-232; clipping bounds of batchglm ``param_bounds``.
* initialize()     diffxpy/testing/tests.py:222-229 -> batchglm ``init_par``: init_a AUTO ->
                   "closed_form" for 0/1 designs (lstsq of log group means through the unique design
                   rows) else "standard" (intercept = log mean); init_b AUTO -> "standard"
                   (intercept = log method-of-moments size).
* train_sequence() diffxpy/testing/tests.py:242-245 -> batchglm numpy ``EstimatorGlm.train`` with the
                   "DEFAULT" strategy {max_steps 1000, update_b_freq 5, method_b brent, ftol_b 1e-6,
                   max_iter_b 1000}: IWLS (Fisher scoring) steps for the location model, a per-gene
                   1-D scale optimisation every ``update_b_freq``-th step, per-gene reverts of steps
                   that worsen the likelihood, convergence on the relative likelihood change 1e-10.
* finalize()       diffxpy/testing/tests.py:247-248: fisher_inv = inv(X^T W X) (location block),
                   log_likelihood, jacobian = sum_k |d ll / d theta_k| / N.

Two scale-step modes:

``method_b="brent"``   reference-faithful: scipy.optimize.brent on -ll over b, bracket (b-20, b+20),
                       tol=ftol_b; naive batchglm likelihood arithmetic; np.linalg.cond/solve guard.
                       Only defined for a single scale coefficient (as in the numpy backend).
``method_b="newton"``  the algorithm the CUDA path implements, restated on the CPU: safeguarded Newton
                       in b with digamma/trigamma, cancellation-free likelihood, Cholesky pivot guard.
                       Works for any number of scale coefficients.  Convergence flags / niter of the
                       CUDA path are compared bit-for-bit against this mode.
"""
import sys
import numpy as np
import scipy.linalg
import scipy.optimize
import scipy.special

ACCURACY_MARGIN_RELATIVE_TO_LIMIT = 2.5   # batchglm pkg_constants
LLTOL_BY_FEATURE = 1e-10                  # batchglm pkg_constants
TINY = np.nextafter(0.0, 1.0)

# error / status bits (shared with include/dxb200.h)
FLAG_CONVERGED = 1          # fully converged before max_steps
FLAG_SINGULAR_FIM = 2       # at least one IWLS step was skipped because X^T W X was singular
FLAG_SCALE_FROZEN = 4       # scale linear predictor sits on the numerical plateau (Poisson limit)
FLAG_NONFINITE = 8          # a non-finite likelihood was met
FLAG_ALL_ZERO = 16          # gene has no non-zero observation

ETA_SCALE_FROZEN = 34.5     # exp(34.5) ~ 1e15: beyond this the NB is numerically Poisson
NEWTON_MAX_ITER = 100
NEWTON_XTOL = 1e-8
NEWTON_MAX_STEP = 5.0
NEWTON_MAX_HALVINGS = 8
NEWTON_GAIN_RTOL = 1e-14
CHOL_PIVOT_RTOL = 2.0 ** -50


def param_bounds():
    """batchglm ``param_bounds`` for float64 (models/base_glm): log-space limits / 2.5."""
    dmax = np.finfo(np.float64).max
    sf = ACCURACY_MARGIN_RELATIVE_TO_LIMIT
    lo = np.log(TINY) / sf
    hi = np.nextafter(np.log(dmax), -np.inf) / sf
    return lo, hi


ETA_MIN, ETA_MAX = param_bounds()
LL_MIN = np.log(TINY)


def _lgamma_ratio(r, y):
    """G(r, y) = lgamma(r + y) - lgamma(r) - y log r, stable for large r."""
    r = np.broadcast_to(r, np.broadcast(r, y).shape).astype(np.float64)
    y = np.broadcast_to(y, r.shape).astype(np.float64)
    out = np.zeros(r.shape)
    big = r >= 1e7
    small = ~big
    if np.any(small):
        rs, ys = r[small], y[small]
        out[small] = scipy.special.gammaln(rs + ys) - scipy.special.gammaln(rs) - ys * np.log(rs)
    if np.any(big):
        rb, yb = r[big], y[big]
        out[big] = yb * (yb - 1.0) / (2.0 * rb) - yb * (yb - 1.0) * (2.0 * yb - 1.0) / (12.0 * rb * rb)
    return out


class NBModel:
    """Dense (cells x genes) NB GLM model terms, batchglm numpy ``glm_nb`` Model."""

    def __init__(self, x, design_loc, design_scale, constraints_loc=None, constraints_scale=None,
                 size_factors=None, stable=False):
        self.x = np.asarray(x, dtype=np.float64)
        self.design_loc = np.asarray(design_loc, dtype=np.float64)
        self.design_scale = np.asarray(design_scale, dtype=np.float64)
        self.constraints_loc = np.eye(self.design_loc.shape[1]) if constraints_loc is None \
            else np.asarray(constraints_loc, dtype=np.float64)
        self.constraints_scale = np.eye(self.design_scale.shape[1]) if constraints_scale is None \
            else np.asarray(constraints_scale, dtype=np.float64)
        self.xh_loc = self.design_loc @ self.constraints_loc
        self.xh_scale = self.design_scale @ self.constraints_scale
        self.log_sf = None if size_factors is None else np.log(np.asarray(size_factors, dtype=np.float64))
        self.stable = stable
        self.lgam_y1 = scipy.special.gammaln(self.x + 1.0)

    # -- linear predictors -------------------------------------------------------------------
    def eta_loc(self, a, j):
        eta = self.xh_loc @ a[:, j]
        if self.log_sf is not None:
            eta = eta + self.log_sf[:, None]
        return np.clip(eta, ETA_MIN, ETA_MAX)

    def eta_scale(self, b, j):
        return np.clip(self.xh_scale @ b[:, j], ETA_MIN, ETA_MAX)

    # -- likelihood ----------------------------------------------------------------------------
    def ll_elem(self, eta_loc, eta_scale, j):
        x = self.x[:, j]
        loc = np.exp(eta_loc)
        scale = np.exp(eta_scale)
        if self.stable:
            l1p = np.log1p(loc / scale)
            ll = _lgamma_ratio(scale, x) - self.lgam_y1[:, j] + x * (eta_loc - l1p) - scale * l1p
            return ll
        # batchglm glm_nb ``ll``: naive arithmetic, elementwise clip to [log(tiny), 0]
        log_r_plus_mu = np.log(scale + loc)
        ll = scipy.special.gammaln(scale + x) - self.lgam_y1[:, j] - scipy.special.gammaln(scale) + \
            x * (eta_loc - log_r_plus_mu) + scale * (eta_scale - log_r_plus_mu)
        return np.clip(ll, LL_MIN, 0.0)

    def ll_byfeature(self, a, b, j):
        return np.sum(self.ll_elem(self.eta_loc(a, j), self.eta_scale(b, j), j), axis=0)

    # -- location model: IWLS terms --------------------------------------------------------------
    def fim_aa_and_jac_a(self, a, b, j):
        """A = Xh^T diag(W) Xh (G x p x p), jac_a = Xh^T (W * ybar) (G x p); W = mu r / (r + mu)."""
        loc = np.exp(self.eta_loc(a, j))
        scale = np.exp(self.eta_scale(b, j))
        w = loc * scale / (scale + loc)
        ybar = (self.x[:, j] - loc) / loc
        xhw = np.einsum('ob,of->fob', self.xh_loc, w)
        fim = np.einsum('fob,oc->fbc', xhw, self.xh_loc)
        jac = np.einsum('fob,of->fb', xhw, ybar)
        return fim, jac

    # -- scale model: gradient / curvature wrt eta_scale -----------------------------------------
    def scale_grad_hess(self, a, b, j):
        """Per gene gradient (G x ps) and Hessian (G x ps x ps) of ll wrt b."""
        x = self.x[:, j]
        eta_loc = self.eta_loc(a, j)
        eta_scale = self.eta_scale(b, j)
        loc = np.exp(eta_loc)
        r = np.exp(eta_scale)
        dpsi = np.where(x > 0, scipy.special.digamma(r + x) - scipy.special.digamma(r), 0.0)
        dpsi1 = np.where(x > 0, scipy.special.polygamma(1, r + x) - scipy.special.polygamma(1, r), 0.0)
        g = r * (dpsi - np.log1p(loc / r) + (loc - x) / (r + loc))
        h = g + r * r * dpsi1 + r * (loc * loc + r * x) / ((r + loc) ** 2)
        grad = np.einsum('ob,of->fb', self.xh_scale, g)
        hess = np.einsum('fob,oc->fbc', np.einsum('ob,of->fob', self.xh_scale, h), self.xh_scale)
        return grad, hess


# ------------------------------------------------------------------------------------------------
# initialisation (batchglm init_par / closedform_nb_glm_logmu / closedform_nb_glm_logphi)
# ------------------------------------------------------------------------------------------------
def init_par(x, design_loc, design_scale, constraints_loc, constraints_scale, size_factors,
             init_a="AUTO", init_b="AUTO"):
    x = np.asarray(x, dtype=np.float64)
    n, g = x.shape
    xs = x if size_factors is None else x / np.asarray(size_factors, dtype=np.float64)[:, None]
    p_loc = constraints_loc.shape[1]
    p_scale = constraints_scale.shape[1]
    train_loc = True
    if isinstance(init_a, str):
        mode = init_a.lower()
        if mode == "auto":
            vals = np.unique(design_loc)
            one_hot = len(vals) == 2 and vals.min() == 0.0 and vals.max() == 1.0
            mode = "closed_form" if one_hot else "standard"
        if mode == "closed_form":
            uniq, inv = np.unique(design_loc, axis=0, return_inverse=True)
            inv = np.asarray(inv).reshape(-1)
            if uniq.shape[0] > 500:
                raise ValueError("large least-square problem in init, likely defined a numeric predictor "
                                 "as categorical")
            means = np.stack([xs[inv == k].mean(axis=0) for k in range(uniq.shape[0])], axis=0)
            linked = np.log(means + TINY)
            ud = uniq @ constraints_loc
            a0, rmsd, rank, _ = np.linalg.lstsq(ud, linked, rcond=None)
            train_loc = not (rmsd.size == 0 or np.all(np.abs(rmsd) < 1e-20))
            if size_factors is not None and np.any(np.asarray(size_factors) != 1):
                train_loc = True
        elif mode == "standard":
            a0 = np.zeros((p_loc, g))
            with np.errstate(divide="ignore"):
                a0[0, :] = np.log(np.mean(x, axis=0))
        elif mode == "all_zero":
            a0 = np.zeros((p_loc, g))
        else:
            raise ValueError("init_a string %s not recognized" % init_a)
    else:
        a0 = np.array(init_a, dtype=np.float64)
    if isinstance(init_b, str):
        mode = init_b.lower()
        if mode == "auto":
            mode = "standard"
        if mode == "standard":
            m = xs.mean(axis=0)
            v = (xs * xs).mean(axis=0) - m * m
            denom = np.fmax(v - m, np.sqrt(TINY))
            b0 = np.zeros((p_scale, g))
            b0[0, :] = np.log(m * m / denom + TINY)
        elif mode == "closed_form":
            uniq, inv = np.unique(design_scale, axis=0, return_inverse=True)
            inv = np.asarray(inv).reshape(-1)
            m = np.stack([xs[inv == k].mean(axis=0) for k in range(uniq.shape[0])], axis=0)
            v = np.stack([(xs[inv == k] ** 2).mean(axis=0) for k in range(uniq.shape[0])], axis=0) - m * m
            denom = np.fmax(v - m, np.sqrt(TINY))
            with np.errstate(divide="ignore"):
                linked = np.log(m * m / denom)
            b0 = np.linalg.lstsq(uniq @ constraints_scale, linked, rcond=None)[0]
        elif mode == "all_zero":
            b0 = np.zeros((p_scale, g))
        else:
            raise ValueError("init_b string %s not recognized" % init_b)
    else:
        b0 = np.array(init_b, dtype=np.float64)
    a0 = np.clip(a0, ETA_MIN, ETA_MAX)
    b0 = np.clip(b0, ETA_MIN, ETA_MAX)
    return a0, b0, train_loc


# ------------------------------------------------------------------------------------------------
# small linear algebra with the guards of the two modes
# ------------------------------------------------------------------------------------------------
def chol_guarded(m):
    """Lower Cholesky factor of an SPD matrix or None if a pivot fails the relative guard
    (v_j <= CHOL_PIVOT_RTOL * m_jj or non-finite).  Same recurrence as the CUDA warp routine."""
    p = m.shape[0]
    l = np.zeros_like(m)
    for j in range(p):
        v = m[j, j] - np.dot(l[j, :j], l[j, :j])
        if not np.isfinite(v) or not (v > CHOL_PIVOT_RTOL * m[j, j]) or not (m[j, j] > 0):
            return None
        d = np.sqrt(v)
        l[j, j] = d
        for i in range(j + 1, p):
            l[i, j] = (m[i, j] - np.dot(l[i, :j], l[j, :j])) / d
    return l


def chol_solve(l, rhs):
    y = scipy.linalg.solve_triangular(l, rhs, lower=True)
    return scipy.linalg.solve_triangular(l.T, y, lower=False)


# ------------------------------------------------------------------------------------------------
# the estimator
# ------------------------------------------------------------------------------------------------
def fit_nb(x, design_loc, design_scale, constraints_loc=None, constraints_scale=None, size_factors=None,
           init_a="AUTO", init_b="AUTO", quick_scale=False, method_b="brent",
           max_steps=1000, update_b_freq=5, ftol_b=1e-6, max_iter_b=1000, lltol=LLTOL_BY_FEATURE,
           verbose=False):
    """Fit the NB GLM for every gene (column of x) and finalize.  Returns a dict with
    a_var (p_loc x G), b_var (p_scale x G), fisher_inv (G x p_loc x p_loc), log_likelihood (G),
    jacobian (G), niter (G, int32), error_codes (G, int32 flag bits), train_loc (bool)."""
    x = np.asarray(x, dtype=np.float64)
    n, g = x.shape
    faithful = method_b == "brent"
    model = NBModel(x, design_loc, design_scale, constraints_loc, constraints_scale, size_factors,
                    stable=not faithful)
    p_loc = model.xh_loc.shape[1]
    p_scale = model.xh_scale.shape[1]
    if faithful and p_scale != 1:
        raise ValueError("method_b='brent' (batchglm numpy backend) supports a single scale coefficient")
    if np.linalg.matrix_rank(model.xh_loc) < p_loc:
        raise ValueError("design_loc @ constraints_loc is not full rank")
    if np.linalg.matrix_rank(model.xh_scale) < p_scale:
        raise ValueError("design_scale @ constraints_scale is not full rank")
    a, b, train_loc = init_par(x, model.design_loc, model.design_scale, model.constraints_loc,
                               model.constraints_scale, size_factors, init_a, init_b)
    train_scale = not quick_scale
    flags = np.zeros(g, dtype=np.int32)
    flags[np.asarray((x != 0).sum(axis=0) == 0)] |= FLAG_ALL_ZERO
    niter = np.zeros(g, dtype=np.int32)

    if train_scale:
        freq = update_b_freq if train_loc else 1
    else:
        freq = np.inf
    epochs_until_b = freq
    allg = np.arange(g)
    loc_conv = np.zeros(g, dtype=bool)
    fully = np.zeros(g, dtype=bool)
    nll = -model.ll_byfeature(a, b, allg)
    step = 0
    while np.any(~fully) and step < max_steps:
        nll_new = nll.copy()
        b_updated = False
        if epochs_until_b == 0:
            idx = np.where(~fully)[0]
            if train_scale and len(idx):
                b_prop = b.copy()
                if faithful:
                    b_prop[:, idx] = _b_step_brent(model, a, b, idx, ftol_b, max_iter_b)
                else:
                    b_prop[:, idx] = _b_step_newton(model, a, b, idx, flags)
                nll_prop = -model.ll_byfeature(a, b_prop, idx)
                bad = nll_prop > nll[idx]
                if not faithful:
                    bad = bad | ~np.isfinite(nll_prop)
                ok = ~bad
                b[:, idx[ok]] = b_prop[:, idx[ok]]
                nll_new[idx[ok]] = nll_prop[ok]
                flags[idx[~np.isfinite(nll_prop)]] |= FLAG_NONFINITE
            epochs_until_b = freq
            b_updated = True
        else:
            idx = np.where(~loc_conv & ~fully)[0]
            if train_loc and len(idx):
                delta = _iwls_step(model, a, b, idx, flags, faithful)
                a_prop = a.copy()
                a_prop[:, idx] = np.clip(a[:, idx] + delta, ETA_MIN, ETA_MAX)
                nll_prop = -model.ll_byfeature(a_prop, b, idx)
                bad = nll_prop > nll[idx]
                if not faithful:
                    bad = bad | ~np.isfinite(nll_prop)
                ok = ~bad
                a[:, idx[ok]] = a_prop[:, idx[ok]]
                nll_new[idx[ok]] = nll_prop[ok]
                flags[idx[~np.isfinite(nll_prop)]] |= FLAG_NONFINITE
            epochs_until_b -= 1
        with np.errstate(divide="ignore", invalid="ignore"):
            conv_f = (nll - nll_new) / nll < lltol
        nll = nll_new
        active = ~fully
        niter[active] += 1
        if b_updated:
            newly = active & loc_conv & conv_f
            fully = fully | newly
            loc_conv = fully.copy()
        else:
            loc_conv = loc_conv | (active & conv_f)
            if not train_scale:
                fully = fully | loc_conv
        step += 1
        if verbose:
            sys.stdout.write("iter %3d: nll=%.6f, converged %d/%d\n" % (step, float(np.sum(nll)), int(fully.sum()), g))
    flags[fully] |= FLAG_CONVERGED
    # frozen scale plateau flag (information only)
    frozen = np.max(model.eta_scale(b, allg), axis=0) >= ETA_SCALE_FROZEN
    flags[frozen] |= FLAG_SCALE_FROZEN

    # finalize ------------------------------------------------------------------------------------
    fim, jac_a = model.fim_aa_and_jac_a(a, b, allg)
    fisher_inv = np.full((g, p_loc, p_loc), np.nan)
    for i in range(g):
        if faithful:
            try:
                fisher_inv[i] = np.linalg.inv(fim[i])
            except np.linalg.LinAlgError:
                pass
        else:
            l = chol_guarded(fim[i])
            if l is not None:
                fisher_inv[i] = chol_solve(l, np.eye(p_loc))
                fisher_inv[i] = 0.5 * (fisher_inv[i] + fisher_inv[i].T)
    grad_b, _ = model.scale_grad_hess(a, b, allg)
    jacobian = (np.sum(np.abs(jac_a), axis=1) + np.sum(np.abs(grad_b), axis=1)) / n
    return dict(a_var=a, b_var=b, fisher_inv=fisher_inv, fim=fim, log_likelihood=-nll,
                jacobian=jacobian, niter=niter, error_codes=flags, train_loc=train_loc,
                jac_a=jac_a, jac_b=grad_b)


def _iwls_step(model, a, b, idx, flags, faithful):
    fim, jac = model.fim_aa_and_jac_a(a, b, idx)
    delta = np.zeros((a.shape[0], len(idx)))
    for t, gi in enumerate(idx):
        if faithful:
            # batchglm iwls_step: solve only where cond(A) < 1/eps
            with np.errstate(all="ignore"):
                try:
                    c = np.linalg.cond(fim[t])
                except np.linalg.LinAlgError:
                    c = np.inf
            if np.isfinite(c) and c < 1.0 / sys.float_info.epsilon:
                delta[:, t] = np.linalg.solve(fim[t], jac[t])
            else:
                flags[gi] |= FLAG_SINGULAR_FIM
        else:
            l = chol_guarded(fim[t])
            if l is not None and np.all(np.isfinite(jac[t])):
                delta[:, t] = chol_solve(l, jac[t])
            else:
                flags[gi] |= FLAG_SINGULAR_FIM
    return delta


def _b_step_brent(model, a, b, idx, ftol, max_iter):
    """batchglm ``_b_step_loop`` / ``optim_handle``: scipy brent on the negative log-likelihood."""
    out = np.zeros((1, len(idx)))
    for t, gi in enumerate(idx):
        j = np.array([gi])
        eta_loc = model.eta_loc(a, j)
        b_j = b[0, gi]
        lb_bracket = max(ETA_MIN, b_j - 20)
        ub_bracket = min(ETA_MAX, b_j + 20)

        def cost(v):
            bb = np.clip(np.array([[v]]), ETA_MIN, ETA_MAX)
            eta_scale = np.clip(model.xh_scale @ bb, ETA_MIN, ETA_MAX)
            return -np.sum(model.ll_elem(eta_loc, eta_scale, j))

        try:
            out[0, t] = scipy.optimize.brent(cost, brack=(lb_bracket, ub_bracket), tol=ftol, maxiter=max_iter)
        except Exception:
            out[0, t] = b_j
    return np.clip(out, ETA_MIN, ETA_MAX)


def _b_step_newton(model, a, b, idx, flags):
    """Safeguarded Newton in b (the CUDA path's scale step, restated)."""
    ps = b.shape[0]
    out = b[:, idx].copy()
    for t, gi in enumerate(idx):
        j = np.array([gi])
        bj = b[:, [gi]].copy()
        if np.max(model.eta_scale(bj, np.array([0]))) >= ETA_SCALE_FROZEN:
            continue
        eta_loc = model.eta_loc(a, j)

        def ll_at(bb):
            return float(np.sum(model.ll_elem(eta_loc, np.clip(model.xh_scale @ bb, ETA_MIN, ETA_MAX), j)))

        f = ll_at(bj)

        for _ in range(NEWTON_MAX_ITER):
            bfull = b.copy()
            bfull[:, [gi]] = bj
            grad, hess = model.scale_grad_hess(a, bfull, j)
            gvec = grad[0]
            hmat = -hess[0]
            if not (np.all(np.isfinite(gvec)) and np.all(np.isfinite(hmat))):
                break
            s = _newton_direction(hmat, gvec)
            smax = np.max(np.abs(s))
            if smax > NEWTON_MAX_STEP:
                s = s * (NEWTON_MAX_STEP / smax)
                smax = NEWTON_MAX_STEP
            pred = float(np.dot(gvec, s))
            # converged: the step is below xtol or its predicted gain is below the resolution of the likelihood
            if smax <= NEWTON_XTOL or not (pred > NEWTON_GAIN_RTOL * abs(f)):
                break
            tstep = 1.0
            accepted = False
            for _h in range(NEWTON_MAX_HALVINGS):
                b_try = np.clip(bj + tstep * s[:, None], ETA_MIN, ETA_MAX)
                f_try = ll_at(b_try)
                if np.isfinite(f_try) and f_try >= f:
                    accepted = True
                    break
                tstep *= 0.5
            if not accepted:
                break
            bj = b_try
            f = f_try
            if np.max(np.clip(model.xh_scale @ bj, ETA_MIN, ETA_MAX)) >= ETA_SCALE_FROZEN:
                break
        out[:, t] = bj[:, 0]
    return out


def _newton_direction(hmat, gvec):
    """Solve (hmat + lam I) s = gvec with hmat = -Hessian; lam grows until Cholesky succeeds."""
    ps = hmat.shape[0]
    if ps == 1:
        if hmat[0, 0] > 0:
            return gvec / hmat[0, 0]
        return np.sign(gvec) * 1.0
    lam = 0.0
    base = max(np.max(np.abs(np.diag(hmat))), 1e-300)
    for _ in range(12):
        l = chol_guarded(hmat + lam * np.eye(ps))
        if l is not None:
            return chol_solve(l, gvec)
        lam = base * 1e-6 if lam == 0.0 else lam * 10.0
    return np.sign(gvec) * 1.0
