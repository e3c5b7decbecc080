"""
Result objects of the single tests (/root/reference/diffxpy/testing/det.py): same classes, lazy ``pval`` /
``qval`` / ``mean`` / ``log_likelihood`` properties, ``summary()`` columns and threshold filters as the
reference (det.py:24-205, 422-458, 461-679, 682-852, 1540-1664, 1667-1816).  Plotting is not part of the
accelerated path and is left out.  All per-gene arithmetic is done by the CUDA kernels.
"""
import abc
import logging
from typing import Union, Dict, Tuple, List, Set

import numpy as np
import pandas as pd
import scipy.sparse
import torch

from .. import _lib
from ..device import CountShard, ptr, stream_ptr
from ..estimator import InputDataGLM
from ..stats import stats
from . import correction
from .utils import dmat_unique

logger = logging.getLogger("diffxpy")


class _DifferentialExpressionTest(metaclass=abc.ABCMeta):
    """Base class (det.py:24-205)."""

    def __init__(self):
        self._pval = None
        self._qval = None
        self._mean = None
        self._log_likelihood = None

    @property
    @abc.abstractmethod
    def gene_ids(self) -> np.ndarray:
        pass

    @property
    @abc.abstractmethod
    def x(self):
        pass

    @abc.abstractmethod
    def log_fold_change(self, base=np.e, **kwargs):
        pass

    def log2_fold_change(self, **kwargs):
        return self.log_fold_change(base=2, **kwargs)

    def log10_fold_change(self, **kwargs):
        return self.log_fold_change(base=10, **kwargs)

    def _test(self, **kwargs) -> np.ndarray:
        pass

    def _correction(self, method) -> np.ndarray:
        """det.py:75-86"""
        if np.all(np.isnan(self.pval)):
            return self.pval
        return correction.correct(pvals=self.pval, method=method)

    def _ave(self):
        pass

    def _ll(self):
        return None

    @property
    def log_likelihood(self):
        if self._log_likelihood is None:
            self._log_likelihood = self._ll()
        return self._log_likelihood

    @property
    def mean(self):
        if self._mean is None:
            self._mean = self._ave()
        return self._mean

    @property
    def pval(self):
        if self._pval is None:
            self._pval = self._test().copy()
        return self._pval

    @property
    def qval(self, method="fdr_bh"):
        if self._qval is None:
            self._qval = self._correction(method=method).copy()
        return self._qval

    @staticmethod
    def _log10_floor(values, floor):
        """log10 of p- or q-values for plotting: zeros and underflows are lifted to the smallest positive double first,
        NaNs become 1 before the clip (so they end at 0 like a p-value of one), and nothing goes below ``floor``
        (behaviour of det.py:120-160)."""
        v = np.asarray(values, dtype=float).reshape(-1)
        out = np.log10(np.maximum(v, np.nextafter(0, 1)))          # (NaN stays NaN through maximum)
        out = np.where(np.isnan(out), 1.0, out)
        return np.minimum(np.maximum(out, floor), 0.0)

    def log10_pval_clean(self, log10_threshold=-30):
        return self._log10_floor(self.pval, log10_threshold)

    def log10_qval_clean(self, log10_threshold=-30):
        return self._log10_floor(self.qval, log10_threshold)

    @abc.abstractmethod
    def summary(self, **kwargs) -> pd.DataFrame:
        pass

    def _threshold_summary(self, res: pd.DataFrame, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None,
                           mean_thres=None) -> pd.DataFrame:
        """Row filter of ``summary()`` (rules of det.py:166-205): q-value at most ``qval_thres`` (NaN q-values drop out),
        fold change outside (fc_lower_thres, fc_upper_thres) on the log2 scale of the table, mean at least ``mean_thres``.
        All given filters must hold."""
        for name, thres in (("fc_lower_thres", fc_lower_thres), ("fc_upper_thres", fc_upper_thres)):
            assert thres is None or thres > 0, "supply positive %s" % name
        keep = np.ones(res.shape[0], dtype=bool)
        if qval_thres is not None:
            with np.errstate(invalid="ignore"):
                keep &= res["qval"].values <= qval_thres          # (False for NaN)
        log2fc = res["log2fc"].values
        outside = []
        if fc_lower_thres is not None:
            outside.append(log2fc <= np.log2(fc_lower_thres))
        if fc_upper_thres is not None:
            outside.append(log2fc >= np.log2(fc_upper_thres))
        if outside:
            keep &= np.logical_or.reduce(outside)
        if mean_thres is not None:
            keep &= res["mean"].values >= mean_thres
        return res.iloc[keep, :]


class _DifferentialExpressionTestSingle(_DifferentialExpressionTest, metaclass=abc.ABCMeta):
    """One p-value and one fold change per gene (det.py:422-458)."""

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        assert self.gene_ids is not None
        res = pd.DataFrame({
            "gene": self.gene_ids,
            "pval": self.pval,
            "qval": self.qval,
            "log2fc": self.log2_fold_change(),
            "mean": self.mean,
            "zero_mean": self.mean == 0
        })
        return res


def _estimator_mean(estim):
    if hasattr(estim, "gene_means"):
        return estim.gene_means()
    return np.asarray(np.mean(estim.x, axis=0)).flatten()


class DifferentialExpressionTestLRT(_DifferentialExpressionTestSingle):
    """Single log-likelihood ratio test per gene (det.py:461-679)."""

    def __init__(self, sample_description: pd.DataFrame, full_design_loc_info, full_estim, reduced_design_loc_info,
                 reduced_estim):
        super().__init__()
        self.sample_description = sample_description
        self.full_design_loc_info = full_design_loc_info
        self.full_estim = full_estim
        self.reduced_design_loc_info = reduced_design_loc_info
        self.reduced_estim = reduced_estim
        self._log_pval = None

    @property
    def gene_ids(self) -> np.ndarray:
        return np.asarray(self.full_estim.input_data.features)

    @property
    def x(self):
        return self.full_estim.x

    @property
    def reduced_model_gradient(self):
        return self.reduced_estim.jacobian

    @property
    def full_model_gradient(self):
        return self.full_estim.jacobian

    def _test(self):
        """det.py:503-514"""
        if np.any(self.full_estim.log_likelihood < self.reduced_estim.log_likelihood):
            logger.warning("Test assumption failed: full model is (partially) less probable than reduced model")
        pval, self._log_pval = stats.likelihood_ratio_test_full(
            ll_full=self.full_estim.log_likelihood,
            ll_reduced=self.reduced_estim.log_likelihood,
            df_full=self.full_estim.input_data.constraints_loc.shape[1] +
            self.full_estim.input_data.constraints_scale.shape[1],
            df_reduced=self.reduced_estim.input_data.constraints_loc.shape[1] +
            self.reduced_estim.input_data.constraints_scale.shape[1],
        )
        return pval

    @property
    def log_pval(self):
        """Accurate natural-log p-value from the chi^2 survival function (no 1 - cdf cancellation)."""
        if self._log_pval is None:
            _ = self.pval
        return self._log_pval

    def _ll(self):
        return self.full_estim.log_likelihood

    def _ave(self):
        return _estimator_mean(self.full_estim)

    def _log_fold_change(self, factors: Union[Dict, Tuple, Set, List], base=np.e):
        """det.py:525-571: pairwise log fold changes between the unique groups spanned by ``factors``."""
        if not (isinstance(factors, list) or isinstance(factors, tuple) or isinstance(factors, set)):
            factors = {factors}
        if not isinstance(factors, set):
            factors = set(factors)
        di = self.full_design_loc_info
        cols_sd = [t for t in di.term_names if t in factors and t in self.sample_description.columns]
        sample_description = self.sample_description[cols_sd]
        dmat = np.array(self.full_estim.input_data.design_loc)
        dmat, sample_description = dmat_unique(dmat, sample_description)
        cols = np.arange(len(di.column_names))
        sel = np.concatenate([cols[di.slice(f)] for f in factors], axis=0)
        neg_sel = np.ones_like(cols).astype(bool)
        neg_sel[sel] = False
        dmat[:, neg_sel] = 0
        dmat, sample_description = dmat_unique(dmat, sample_description)
        locations = self.full_estim.model.inverse_link_loc(np.matmul(dmat, self.full_estim.model.a))
        locations = np.log(locations) / np.log(base)
        dist = np.expand_dims(locations, axis=0)
        dist = np.transpose(dist, [1, 0, 2]) - dist
        return dist

    def log_fold_change(self, base=np.e, return_type="vector"):
        """det.py:573-604"""
        factors = set(self.full_design_loc_info.term_names) - set(self.reduced_design_loc_info.term_names)
        if return_type == "vector":
            if len(factors) > 1 or self.sample_description[list(factors)].drop_duplicates().shape[0] != 2:
                return None
            dists = self._log_fold_change(factors=factors, base=base)
            return dists[1, 0]
        return self._log_fold_change(factors=factors, base=base)

    def _per_category(self, dmat, coefficients, inverse_link):
        """Fitted parameter for every distinct combination of the factors of the full location model (det.py:606-648)."""
        di = self.full_design_loc_info
        factors = []
        for t in di.term_names:
            for f in t.split(":"):
                f = f[2:-1] if f.startswith("C(") and f.endswith(")") else f
                if f in self.sample_description.columns and f not in factors:
                    factors.append(f)
        sample_description = self.sample_description[factors]
        dmat, sample_description = dmat_unique(np.array(dmat), sample_description)
        retval = pd.DataFrame(inverse_link(np.matmul(dmat, coefficients)), columns=self.full_estim.input_data.features)
        for col in sample_description:
            retval[col] = sample_description[col].values
        return retval.set_index(list(sample_description.columns))

    def locations(self):
        """det.py:606-626: pandas.DataFrame of the fitted means for the different categories of the factors."""
        return self._per_category(self.full_estim.input_data.design_loc, self.full_estim.model.a,
                                  self.full_estim.model.inverse_link_loc)

    def scales(self):
        """det.py:628-648: the same for the dispersion model.  (The reference's body cannot run -- ``dmat.doc``,
        ``par_link_scale`` -- this is what it evidently means.)"""
        return self._per_category(self.full_estim.input_data.design_scale, self.full_estim.model.b,
                                  self.full_estim.model.inverse_link_scale)

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        """det.py:650-679"""
        res = super().summary(**kwargs)
        res["grad"] = self.full_model_gradient
        res["grad_red"] = self.reduced_model_gradient
        res = self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                      fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)
        return res


class DifferentialExpressionTestWald(_DifferentialExpressionTestSingle):
    """Single Wald test per gene (det.py:682-852)."""

    def __init__(self, model_estim, col_indices: np.ndarray, noise_model: str, sample_description: pd.DataFrame):
        super().__init__()
        self.sample_description = sample_description
        self.model_estim = model_estim
        self.coef_loc_totest = col_indices
        self.noise_model = noise_model
        self._store_ols = None
        self._log_pval = None
        try:
            self._error_codes = self.model_estim.error_codes if self.model_estim.error_codes is not None else None
        except Exception:
            self._error_codes = None
        try:
            self._niter = self.model_estim.niter if self.model_estim.niter is not None else None
        except Exception:
            self._niter = None

    @property
    def gene_ids(self) -> np.ndarray:
        return np.asarray(self.model_estim.input_data.features)

    @property
    def x(self):
        return self.model_estim.x

    @property
    def model_gradient(self):
        return self.model_estim.jacobian

    def log_fold_change(self, base=np.e, **kwargs):
        """det.py:742-765: the tested coefficient itself (natural log scale, ``base`` is ignored exactly as in
        the reference); with several tested coefficients the one of largest magnitude."""
        if len(self.coef_loc_totest) == 1:
            return self.model_estim.a_var[self.coef_loc_totest][0]
        idx0 = np.argmax(np.abs(self.model_estim.a_var[self.coef_loc_totest]), axis=0)
        idx1 = np.arange(len(idx0))
        return self.model_estim.a_var[self.coef_loc_totest, :][tuple(idx0), tuple(idx1)]

    def _ll(self):
        return self.model_estim.log_likelihood

    def _ave(self):
        return _estimator_mean(self.model_estim)

    def _test(self):
        """det.py:783-811"""
        self.theta_mle = self.model_estim.a_var[self.coef_loc_totest]
        if len(self.coef_loc_totest) == 1:
            self.theta_mle = self.theta_mle[0]
            self.theta_sd = self.model_estim.fisher_inv[:, self.coef_loc_totest[0], self.coef_loc_totest[0]].copy()
            self.theta_sd = np.nextafter(0, np.inf, out=self.theta_sd, where=self.theta_sd < np.nextafter(0, np.inf))
            self.theta_sd = np.sqrt(self.theta_sd)
            pval, self._log_pval = stats.wald_test_full(theta_mle=self.theta_mle, theta_sd=self.theta_sd, theta0=0)
            return pval
        self.theta_sd = np.diagonal(self.model_estim.fisher_inv, axis1=-2, axis2=-1).copy()
        self.theta_sd = np.nextafter(0, np.inf, out=self.theta_sd, where=self.theta_sd < np.nextafter(0, np.inf))
        self.theta_sd = np.sqrt(self.theta_sd)
        pval, self._log_pval = stats.wald_test_chisq_full(
            theta_mle=self.theta_mle,
            theta_covar=self.model_estim.fisher_inv[:, self.coef_loc_totest, :][:, :, self.coef_loc_totest],
            theta0=0)
        return pval

    @property
    def log_pval(self):
        """Accurate natural-log p-value (survival function; does not saturate at |z| > 8.3 like 1 - cdf)."""
        if self._log_pval is None:
            _ = self.pval
        return self._log_pval

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        """det.py:813-852"""
        res = super().summary(**kwargs)
        res["grad"] = self.model_gradient
        if len(self.theta_mle.shape) == 1:
            res["coef_mle"] = self.theta_mle
        if len(self.theta_sd.shape) == 1:
            res["coef_sd"] = self.theta_sd
        if self.log_likelihood is not None:
            res["ll"] = self.log_likelihood
        if self._error_codes is not None:
            res["err"] = self._error_codes
        if self._niter is not None:
            res["niter"] = self._niter
        res = self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                      fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)
        return res


def two_group_postprocess(out, flags, n0, n1, do_rank, is_logged=False, is_sig_zerovar=True):
    """Host-side rules of the model-free tests on the rows K_twogroup produced (det.py:1564-1620, 1691-1743):
    overall mean, run mask, zero-variance p-values, log fold change of sample 1 over sample 0."""
    mean_x0 = out[:, _lib.TG_MEAN0].copy()
    mean_x1 = out[:, _lib.TG_MEAN1].copy()
    var_x0, var_x1 = out[:, _lib.TG_VAR0], out[:, _lib.TG_VAR1]
    mean = np.asarray(np.average(a=np.vstack([mean_x0, mean_x1]),
                                 weights=np.array([n0 / (n0 + n1), n1 / (n0 + n1)]), axis=0, returned=False)).flatten()
    var_geq_zero = np.logical_or(var_x0 > 0, var_x1 > 0)
    idx_run = np.where(np.logical_and(mean != 0, var_geq_zero))[0]
    pval = np.zeros([out.shape[0]]) + np.nan
    col = _lib.TG_PVAL_RANK if do_rank else _lib.TG_PVAL_T
    if do_rank and np.any(flags[idx_run] & _lib.FLAG_RANK_OVERFLOW):
        raise _lib.DxError("rank test: a gene has more non-count values than the rank workspace holds (stats.RANK_WORK_MAX_BYTES)")
    pval[idx_run] = out[idx_run, col]
    pval[np.where(np.logical_and(np.logical_and(mean_x0 == mean_x1, mean > 0), np.logical_not(var_geq_zero)))[0]] = 1.0
    if is_sig_zerovar:
        pval[np.where(np.logical_and(mean_x0 != mean_x1, np.logical_not(var_geq_zero)))[0]] = 0.0
    res = dict(pval=pval, mean=mean, var_geq_zero=var_geq_zero, mean_x0=mean_x0.copy(), mean_x1=mean_x1.copy(),
               var_x0=var_x0, var_x1=var_x1)
    if is_logged:
        res["logfc"] = mean_x1 - mean_x0
    else:
        tiny = np.nextafter(0, np.inf)
        res["logfc"] = np.log(np.maximum(mean_x1, tiny)) - np.log(np.maximum(mean_x0, tiny))
    return res


class _DifferentialExpressionTestTwoGroup(_DifferentialExpressionTestSingle):
    """Shared constructor of the model-free tests (det.py:1545-1620 and 1672-1743): group split by first
    appearance, group means / ddof-0 variances, run mask, zero-variance rules, log fold change.  The moments,
    the Welch p-value and the Mann-Whitney p-value come from K_ingest + K_twogroup on the device."""

    _do_rank = False

    def __init__(self, data, sample_description: pd.DataFrame, grouping, gene_names, is_logged, is_sig_zerovar: bool = True):
        super().__init__()
        if isinstance(data, InputDataGLM):
            data = data.x
        elif hasattr(data, "X") and not isinstance(data, (np.ndarray, CountShard)) and not scipy.sparse.issparse(data):
            data = data.X
        self._x = data
        self.sample_description = sample_description
        self.grouping = grouping
        self._gene_names = np.asarray(gene_names)

        grouping = np.asarray(grouping)
        groups = pd.Series(grouping).unique()          # first-appearance order (utils.py:114)
        if len(groups) < 2:
            raise ValueError("grouping must contain at least two groups")
        cell_group = np.full(grouping.shape[0], _lib.DX_SKIP_GROUP, dtype=np.uint8)
        cell_group[grouping == groups[0]] = 0
        cell_group[grouping == groups[1]] = 1
        n0 = int(np.sum(cell_group == 0))
        n1 = int(np.sum(cell_group == 1))
        shard = CountShard.from_host(data)
        res = stats.two_group_device(shard, cell_group, n0, n1, do_rank=self._do_rank)
        self._u1x2, self._tie, self._flags = res["u1x2"], res["tie"], res["flags"]
        r = two_group_postprocess(res["out"], res["flags"], n0, n1, self._do_rank, is_logged, is_sig_zerovar)
        self._mean, self._ave_nonzero, self._var_geq_zero = r["mean"], r["mean"] != 0, r["var_geq_zero"]
        self._pval, self._logfc = r["pval"], r["logfc"]
        self._mean_x0, self._mean_x1, self._var_x0, self._var_x1 = r["mean_x0"], r["mean_x1"], r["var_x0"], r["var_x1"]

    @property
    def gene_ids(self) -> np.ndarray:
        return self._gene_names

    @property
    def x(self):
        return self._x

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        """det.py:1636-1664 / 1762-1790"""
        res = super().summary(**kwargs)
        res["zero_variance"] = np.logical_not(self._var_geq_zero)
        res = self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                      fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)
        return res


class DifferentialExpressionTestTT(_DifferentialExpressionTestTwoGroup):
    """Single Welch t-test per gene (det.py:1540-1664)."""
    _do_rank = False

    def log_fold_change(self, base=np.e, **kwargs):
        return self._logfc / np.log(base)


class DifferentialExpressionTestRank(_DifferentialExpressionTestTwoGroup):
    """Single rank test per gene, Mann-Whitney U (det.py:1667-1816)."""
    _do_rank = True

    @property
    def u_statistic(self):
        """U1 of the first-appearing group (scipy's ``statistic``), exact: integer 2*U1 halved."""
        return self._u1x2 / 2.0

    def log_fold_change(self, base=np.e, **kwargs):
        if base == np.e:
            return self._logfc
        return self._logfc / np.log(base)


def __getattr__(name):
    """The reference defines the multi-test classes in det.py (det.py:1819-2084); here they live next to the pairwise
    classes in det_pair.py and are resolved lazily to keep the two modules free of a circular import."""
    if name in ("_DifferentialExpressionTestMulti", "DifferentialExpressionTestVsRest", "DifferentialExpressionTestByPartition"):
        from . import det_pair
        return getattr(det_pair, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
