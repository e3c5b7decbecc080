"""
Result wrappers of ``de.test.continuous_1d`` (/root/reference/diffxpy/testing/det_cont.py:21-593, plots excluded):
p / q values come from the wrapped Wald or LRT test, the fold change is the ratio of the largest to the smallest
FITTED expression value along the continuous covariate, and the fitted curves can be queried per gene.  The
per-gene curve scans (a cells x basis by basis x genes product) run on the GPU in gene chunks.
"""
import logging

import numpy as np
import pandas as pd
import torch

from ..device import cuda_device
from .det import _DifferentialExpressionTestSingle

logger = logging.getLogger("diffxpy")


class _DifferentialExpressionTestCont(_DifferentialExpressionTestSingle):
    """det_cont.py:21-547"""

    def __init__(self, de_test, model_estim, size_factors, continuous_coords, spline_coefs, interpolated_spline_basis,
                 noise_model):
        self._de_test = de_test
        self._model_estim = model_estim
        self._size_factors = size_factors
        self._continuous_coords = np.asarray(continuous_coords)
        self._spline_coefs = list(spline_coefs)
        self._interpolated_spline_basis = interpolated_spline_basis
        self.noise_model = noise_model
        self._scan_cache = {}

    @property
    def gene_ids(self) -> np.ndarray:
        return self._de_test.gene_ids

    @property
    def x(self):
        return self._de_test.x

    @property
    def pval(self) -> np.ndarray:
        return self._de_test.pval

    @property
    def qval(self) -> np.ndarray:
        return self._de_test.qval

    @property
    def mean(self) -> np.ndarray:
        return self._de_test.mean

    @property
    def log_likelihood(self) -> np.ndarray:
        return self._de_test.log_likelihood

    def summary(self, non_numeric=False, qval_thres=None, fc_upper_thres=None, fc_lower_thres=None,
                mean_thres=None) -> pd.DataFrame:
        res = self._de_test.summary()
        res['log2fc'] = self.log2_fold_change(non_numeric=non_numeric)      # fold change of the fitted curve
        return self._threshold_summary(res=res, qval_thres=qval_thres, fc_upper_thres=fc_upper_thres,
                                       fc_lower_thres=fc_lower_thres, mean_thres=mean_thres)

    # ---- gene selection (det_cont.py:163-206) ---------------------------------------------------------
    def _idx_genes(self, genes):
        if not isinstance(genes, list):
            if isinstance(genes, (np.ndarray, pd.Index)):
                genes = genes.tolist()
            else:
                genes = [genes]
        ids = self.gene_ids.tolist()
        n_genes = len(ids)
        if isinstance(genes[0], str):
            found = [g for g in genes if g in ids]
            if len(found) != len(genes):
                logger.info("did not find some genes, omitting")
            return np.array([ids.index(g) for g in found], dtype=np.int64), found
        if isinstance(genes[0], (int, np.integer)):
            found = [int(g) for g in genes if g < n_genes]
            if len(found) != len(genes):
                logger.info("did not find some genes, omitting")
            return np.array(found, dtype=np.int64), [ids[g] for g in found]
        raise ValueError("only string and integer elements allowed in genes")

    def _spline_par_loc_idx(self, intercept=True):
        names = list(self._model_estim.input_data.loc_names)
        idx = [names.index(x) for x in self._spline_coefs]
        if 'Intercept' in names and intercept:
            idx = [names.index('Intercept')] + idx
        return np.array(idx, dtype=np.int64)

    # ---- fitted curves ----------------------------------------------------------------------------------
    def _linear_predictor_inputs(self, non_numeric):
        idata = self._model_estim.input_data
        a = np.asarray(self._model_estim.model.a)
        if non_numeric:
            off = idata.size_factors if self._size_factors is not None else None
            return idata.design_loc, a, off
        idx = self._spline_par_loc_idx(intercept=True)
        return idata.design_loc[:, idx], a[idx, :], None

    def _continuous_model(self, idx, non_numeric=False):
        """Fitted expression at every observed cell for the genes ``idx`` (det_cont.py:208-234)."""
        idx = np.atleast_1d(np.asarray(idx))
        dmat, a, off = self._linear_predictor_inputs(non_numeric)
        eta = dmat @ a[:, idx]
        if off is not None:
            eta = eta + off
        return np.exp(eta)

    def _continuous_interpolation(self, idx):
        """Fitted expression on the uniform grid of the interpolated basis (det_cont.py:236-258)."""
        idx = np.atleast_1d(np.asarray(idx))
        a = np.asarray(self._model_estim.model.a)[self._spline_par_loc_idx(intercept=True), :][:, idx]
        eta_loc = self._interpolated_spline_basis[:, :-1] @ a
        return self._interpolated_spline_basis[:, -1], np.exp(eta_loc)

    def _scan(self, non_numeric=False):
        """min / max / argmin / argmax of the linear predictor over the cells, for ALL genes (cached)."""
        key = bool(non_numeric)
        if key not in self._scan_cache:
            dmat, a, off = self._linear_predictor_inputs(non_numeric)
            dev = cuda_device()
            with torch.cuda.device(dev):
                d = torch.from_numpy(np.ascontiguousarray(dmat, dtype=np.float64)).to(dev)
                o = None if off is None else torch.from_numpy(np.ascontiguousarray(off, dtype=np.float64).reshape(-1, 1)).to(dev)
                g = a.shape[1]
                chunk = max(1, min(g, int(2 ** 27 // max(1, d.shape[0]))))
                lo, hi, alo, ahi = [], [], [], []
                for g0 in range(0, g, chunk):
                    eta = d @ torch.from_numpy(np.ascontiguousarray(a[:, g0:g0 + chunk])).to(dev)
                    if o is not None:
                        eta = eta + o
                    mn, amn = eta.min(dim=0)
                    mx, amx = eta.max(dim=0)
                    lo.append(mn.cpu()); hi.append(mx.cpu()); alo.append(amn.cpu()); ahi.append(amx.cpu())
            self._scan_cache[key] = (torch.cat(lo).numpy(), torch.cat(hi).numpy(), torch.cat(alo).numpy(), torch.cat(ahi).numpy())
        return self._scan_cache[key]

    def min_max(self, genes, non_numeric=False):
        idx, _ = self._idx_genes(genes)
        lo, hi, _, _ = self._scan(non_numeric)
        return np.exp(lo[idx]), np.exp(hi[idx])

    def max(self, genes, non_numeric=False):
        return self.min_max(genes, non_numeric)[1]

    def min(self, genes, non_numeric=False):
        return self.min_max(genes, non_numeric)[0]

    def argmax(self, genes, non_numeric=False):
        idx, _ = self._idx_genes(genes)
        return self._continuous_coords[self._scan(non_numeric)[3][idx]]

    def argmin(self, genes, non_numeric=False):
        idx, _ = self._idx_genes(genes)
        return self._continuous_coords[self._scan(non_numeric)[2][idx]]

    def log_fold_change(self, base=np.e, genes=None, non_numeric=False):
        """log of (largest fitted value / smallest fitted value) along the covariate (det_cont.py:110-131)."""
        if genes is None:
            idx = np.arange(0, len(self.gene_ids))
        else:
            idx, _ = self._idx_genes(genes)
        lo, hi, _, _ = self._scan(non_numeric)
        min_val, max_val = np.exp(lo[idx]), np.exp(hi[idx])
        tiny = np.nextafter(0, 1)
        max_val = np.where(max_val == 0, tiny, max_val)
        min_val = np.where(min_val == 0, tiny, min_val)
        return (np.log(max_val) - np.log(min_val)) / np.log(base)

    def log2_fold_change(self, genes=None, non_numeric=False):
        return self.log_fold_change(base=2, genes=genes, non_numeric=non_numeric)

    def log10_fold_change(self, genes=None, non_numeric=False):
        return self.log_fold_change(base=10, genes=genes, non_numeric=non_numeric)


class DifferentialExpressionTestWaldCont(_DifferentialExpressionTestCont):
    """det_cont.py:550-570"""

    def __init__(self, de_test, size_factors, continuous_coords, spline_coefs, interpolated_spline_basis, noise_model):
        super().__init__(de_test=de_test, model_estim=de_test.model_estim, size_factors=size_factors,
                         continuous_coords=continuous_coords, spline_coefs=spline_coefs,
                         interpolated_spline_basis=interpolated_spline_basis, noise_model=noise_model)


class DifferentialExpressionTestLRTCont(_DifferentialExpressionTestCont):
    """det_cont.py:573-593"""

    def __init__(self, de_test, size_factors, continuous_coords, spline_coefs, interpolated_spline_basis, noise_model):
        super().__init__(de_test=de_test, model_estim=de_test.full_estim, size_factors=size_factors,
                         continuous_coords=continuous_coords, spline_coefs=spline_coefs,
                         interpolated_spline_basis=interpolated_spline_basis, noise_model=noise_model)
