"""
Result objects of the multi-test wrappers (/root/reference/diffxpy/testing/det.py:1819-2084 and det_pair.py):
one p-value / fold change per gene AND per comparison.  ``pval`` has shape (groups, groups, genes) for the
pairwise tests and (1, groups or partitions, genes) for versus-rest / by-partition tests; ``summary()`` reports the
per-gene minimum p / q value and the extreme fold change like the reference (det.py:1846-1887).
"""
import abc
from typing import Union

import numpy as np
import pandas as pd

from ..stats import stats
from . import correction
from .det import _DifferentialExpressionTest, _estimator_mean


class _DifferentialExpressionTestMulti(_DifferentialExpressionTest, metaclass=abc.ABCMeta):
    """det.py:1819-1887"""

    def __init__(self, correction_type: str):
        super().__init__()
        if correction_type.lower() not in ("global", "by_test"):
            raise ValueError("correction_type %s not recognized" % correction_type)
        self._correction_type = correction_type

    def _correction(self, method):
        pval = self.pval
        if self._correction_type.lower() == "global":
            return np.reshape(correction.correct(pvals=np.reshape(pval, -1), method=method), pval.shape)
        flat = pval.reshape(-1, pval.shape[-1])
        return np.stack([correction.correct(pvals=row, method=method) for row in flat]).reshape(pval.shape)

    def summary(self, **kwargs) -> pd.DataFrame:
        assert self.gene_ids is not None
        raw_logfc = self.log_fold_change(base=2.)
        flat_logfc = raw_logfc.reshape(-1, raw_logfc.shape[-1])
        r, c = np.unravel_index(flat_logfc.argmax(0), raw_logfc.shape[:2])
        # a maximum found in the lower triangle belongs to the mirrored comparison: flip its sign
        logfc = raw_logfc[r, c, np.arange(raw_logfc.shape[-1])] * np.where(r <= c, 1, -1)
        return pd.DataFrame({
            "gene": self.gene_ids,
            "pval": np.min(self.pval.reshape(-1, self.pval.shape[-1]), axis=0),
            "qval": np.min(self.qval.reshape(-1, self.qval.shape[-1]), axis=0),
            "log2fc": np.asarray(logfc),
            "mean": np.asarray(self.mean)
        })


# ----------------------------------------------------------------------------------------------------------
# pairwise (det_pair.py)
# ----------------------------------------------------------------------------------------------------------
class _DifferentialExpressionTestPairwiseBase(_DifferentialExpressionTestMulti):
    """det_pair.py:20-262"""
    groups: list

    def _get_group_idx(self, groups0, groups1):
        if np.any([x not in self.groups for x in groups0]):
            raise ValueError('element of groups1 not recognized; not in groups %s' % self.groups)
        if np.any([x not in self.groups for x in groups1]):
            raise ValueError('element of groups2 not recognized; not in groups %s' % self.groups)
        return np.array([self.groups.index(x) for x in groups0]), np.array([self.groups.index(x) for x in groups1])

    def _correction_pairs(self, idx0, idx1, method):
        pval = self._pval_pairs(idx0=idx0, idx1=idx1)
        if self._correction_type.lower() == "global":
            return np.reshape(correction.correct(pvals=np.reshape(pval, -1), method=method), pval.shape)
        flat = pval.reshape(-1, pval.shape[-1])
        return np.stack([correction.correct(pvals=row, method=method) for row in flat]).reshape(pval.shape)

    def log_fold_change(self, base=np.e, **kwargs):
        return self._logfc if base == np.e else self._logfc / np.log(base)

    def _pval_pairs(self, idx0, idx1):
        return self.pval[idx0, :, :][:, idx1, :]

    def _qval_pairs(self, idx0, idx1, method="fdr_bh"):
        return self._correction_pairs(idx0=idx0, idx1=idx1, method=method)

    def _log_fold_change_pairs(self, idx0, idx1, base):
        return self.log_fold_change(base=base)[idx0, :, :][:, idx1, :]

    def pval_pairs(self, groups0, groups1):
        idx0, idx1 = self._get_group_idx(groups0=groups0, groups1=groups1)
        return self._pval_pairs(idx0=idx0, idx1=idx1)

    def qval_pairs(self, groups0, groups1, method="fdr_bh"):
        idx0, idx1 = self._get_group_idx(groups0=groups0, groups1=groups1)
        return self._qval_pairs(idx0=idx0, idx1=idx1, method=method)

    def log10_pval_pairs_clean(self, groups0, groups1, log10_threshold=-30):
        pvals = np.reshape(self.pval_pairs(groups0=groups0, groups1=groups1), -1)
        pvals = np.nextafter(0, 1, out=pvals, where=pvals == 0)
        log10_pval_clean = np.log(pvals) / np.log(10)
        log10_pval_clean[np.isnan(log10_pval_clean)] = 1
        return np.clip(log10_pval_clean, log10_threshold, 0, log10_pval_clean)

    def log10_qval_pairs_clean(self, groups0, groups1, log10_threshold=-30):
        qvals = np.reshape(self.qval_pairs(groups0=groups0, groups1=groups1), -1)
        qvals = np.nextafter(0, 1, out=qvals, where=qvals == 0)
        log10_qval_clean = np.log(qvals) / np.log(10)
        log10_qval_clean[np.isnan(log10_qval_clean)] = 1
        return np.clip(log10_qval_clean, log10_threshold, 0, log10_qval_clean)

    def log_fold_change_pairs(self, groups0, groups1, base=np.e):
        idx0, idx1 = self._get_group_idx(groups0=groups0, groups1=groups1)
        return self._log_fold_change_pairs(idx0=idx0, idx1=idx1, base=base)

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        res = super().summary(**kwargs)
        return self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                       fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)

    def summary_pairs(self, groups0, groups1, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None,
                      mean_thres=None, method="fdr_bh") -> pd.DataFrame:
        """det_pair.py:203-262"""
        assert self.gene_ids is not None
        pval = self.pval_pairs(groups0=groups0, groups1=groups1)
        qval = self.qval_pairs(groups0=groups0, groups1=groups1, method=method)
        raw_logfc = self.log_fold_change_pairs(groups0=groups0, groups1=groups1, base=2)
        flat_logfc = raw_logfc.reshape(-1, raw_logfc.shape[-1])
        r, c = np.unravel_index(flat_logfc.argmax(0), raw_logfc.shape[:2])
        logfc = raw_logfc[r, c, np.arange(raw_logfc.shape[-1])] * np.where(r <= c, 1, -1)
        res = pd.DataFrame({
            "gene": self.gene_ids,
            "pval": np.min(pval, axis=(0, 1)),
            "qval": np.min(qval, axis=(0, 1)),
            "log2fc": np.asarray(logfc),
            "mean": np.asarray(self.mean)
        })
        return self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                       fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)


class DifferentialExpressionTestPairwiseStandard(_DifferentialExpressionTestPairwiseBase):
    """Pairwise tests computed comparison by comparison (det_pair.py:265-334)."""

    def __init__(self, gene_ids, pval, logfc, ave, groups, tests, correction_type: str):
        super().__init__(correction_type=correction_type)
        self._gene_ids = np.asarray(gene_ids)
        self._pval = pval
        self._logfc = logfc
        self._mean = np.asarray(ave).flatten()
        self.groups = list(np.asarray(groups))
        self._tests = tests
        _ = self.qval

    @property
    def gene_ids(self) -> np.ndarray:
        return self._gene_ids

    @property
    def x(self):
        return None

    @property
    def tests(self):
        if self._tests is None:
            raise ValueError("Individual tests were not kept!")
        return self._tests


class DifferentialExpressionTestZTest(_DifferentialExpressionTestPairwiseBase):
    """Pairwise z-tests between the group coefficients of ONE fit '~ 0 + group' (det_pair.py:337-443, 519-622).
    All pairs of all genes go through a single launch of the two-coefficient z-test kernel; the lazy variant
    of the reference (``DifferentialExpressionTestZTestLazy``) is therefore the same object."""

    def __init__(self, model_estim, grouping, groups, correction_type: str):
        super().__init__(correction_type=correction_type)
        self.model_estim = model_estim
        self.grouping = grouping
        self.groups = list(np.asarray(groups))
        self._theta_mle = np.asarray(model_estim.a_var)                                   # groups x genes
        var = np.diagonal(model_estim.fisher_inv, axis1=-2, axis2=-1).T.copy()
        tiny = np.nextafter(0, np.inf)
        var[var < tiny] = tiny                                                             # NaN stays NaN
        self._theta_sd = np.sqrt(var)
        k, g = self._theta_mle.shape
        self._logfc = self._theta_mle[None, :, :] - self._theta_mle[:, None, :]           # [i, j] = theta_j - theta_i
        _ = self.pval
        _ = self.qval

    def _test(self, **kwargs):
        k, g = self._theta_mle.shape
        pvals = np.tile(np.nan, [k, k, g])
        iu, ju = np.triu_indices(k, 1)
        if len(iu):
            p = stats.two_coef_z_test(
                theta_mle0=np.ascontiguousarray(self._theta_mle[iu]).reshape(-1),
                theta_mle1=np.ascontiguousarray(self._theta_mle[ju]).reshape(-1),
                theta_sd0=np.ascontiguousarray(self._theta_sd[iu]).reshape(-1),
                theta_sd1=np.ascontiguousarray(self._theta_sd[ju]).reshape(-1)).reshape(len(iu), g)
            pvals[iu, ju] = p
            pvals[ju, iu] = p
        return pvals

    @property
    def gene_ids(self) -> np.ndarray:
        return np.asarray(self.model_estim.input_data.features)

    @property
    def x(self):
        return self.model_estim.x

    @property
    def log_likelihood(self):
        return self.model_estim.log_likelihood

    @property
    def model_gradient(self):
        return self.model_estim.jacobian

    def _ave(self):
        return _estimator_mean(self.model_estim)


DifferentialExpressionTestZTestLazy = DifferentialExpressionTestZTest
_DifferentialExpressionTestPairwiseLazyBase = _DifferentialExpressionTestPairwiseBase


# ----------------------------------------------------------------------------------------------------------
# versus rest / by partition (det.py:1890-2084)
# ----------------------------------------------------------------------------------------------------------
class DifferentialExpressionTestVsRest(_DifferentialExpressionTestMulti):
    """Each group against the union of all others (det.py:1890-2015)."""

    def __init__(self, gene_ids, pval, logfc, ave, groups, tests, correction_type: str):
        super().__init__(correction_type=correction_type)
        self._gene_ids = np.asarray(gene_ids)
        self._pval = pval
        self._logfc = logfc
        self._mean = np.asarray(ave).flatten()
        self.groups = list(np.asarray(groups))
        self._tests = tests
        _ = self.qval

    @property
    def tests(self):
        if self._tests is None:
            raise ValueError("Individual tests were not kept!")
        return self._tests

    @property
    def gene_ids(self) -> np.ndarray:
        return self._gene_ids

    @property
    def x(self) -> Union[np.ndarray, None]:
        return None

    def log_fold_change(self, base=np.e, **kwargs):
        return self._logfc if base == np.e else self._logfc / np.log(base)

    def _check_group(self, group):
        if group not in self.groups:
            raise ValueError('group "%s" not recognized' % group)

    def pval_group(self, group):
        self._check_group(group)
        return self.pval[0, self.groups.index(group), :]

    def qval_group(self, group):
        self._check_group(group)
        return self.qval[0, self.groups.index(group), :]

    def log_fold_change_group(self, group, base=np.e):
        self._check_group(group)
        return self.log_fold_change(base=base)[0, self.groups.index(group), :]

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        res = super().summary(**kwargs)
        return self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                       fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)

    def summary_group(self, group, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None) -> pd.DataFrame:
        assert self.gene_ids is not None
        res = pd.DataFrame({
            "gene": self.gene_ids,
            "pval": self.pval_group(group=group),
            "qval": self.qval_group(group=group),
            "log2fc": self.log_fold_change_group(group=group, base=2),
            "mean": np.asarray(self.mean)
        })
        return self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                       fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)


class DifferentialExpressionTestByPartition(_DifferentialExpressionTestMulti):
    """One single test per partition of the cells (det.py:2018-2084)."""

    def __init__(self, partitions, tests, ave, correction_type: str = "by_test"):
        super().__init__(correction_type=correction_type)
        self.partitions = list(np.asarray(partitions))
        self._tests = tests
        self._gene_ids = tests[0].gene_ids
        self._pval = np.expand_dims(np.vstack([x.pval for x in tests]), axis=0)
        self._logfc = np.expand_dims(np.vstack([x.log_fold_change() for x in tests]), axis=0)
        self._mean = np.asarray(ave).flatten()
        _ = self.qval

    @property
    def gene_ids(self) -> np.ndarray:
        return self._gene_ids

    @property
    def x(self):
        return None

    def log_fold_change(self, base=np.e, **kwargs):
        return self._logfc if base == np.e else self._logfc / np.log(base)

    def _check_partition(self, partition):
        if partition not in self.partitions:
            raise ValueError('partition "%s" not recognized' % partition)

    @property
    def tests(self):
        if self._tests is None:
            raise ValueError("Individual tests were not kept!")
        return self._tests

    def test_of(self, partition):
        self._check_partition(partition)
        return self._tests[self.partitions.index(partition)]

    def summary(self, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None, mean_thres=None, **kwargs) -> pd.DataFrame:
        res = super().summary(**kwargs)
        return self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                       fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)
