"""
Input parsing helpers (/root/reference/diffxpy/testing/utils.py:20-124) and the relays to the design /
constraint builders (utils.py:127-319).  AnnData is duck-typed (``.X``, ``.obs``, ``.var_names``) because the
package itself is optional.
"""
from typing import List, Tuple, Union

import numpy as np
import pandas as pd
import scipy.sparse

from ..data import design as _design
from ..data.design import constraint_matrix_from_string, constraint_matrix_from_dict, view_coef_names  # noqa: F401
from ..device import CountShard
from ..estimator import InputDataGLM


def _is_anndata_like(data):
    return hasattr(data, "X") and not isinstance(data, (np.ndarray, InputDataGLM, CountShard)) \
        and not scipy.sparse.issparse(data)


def parse_gene_names(data, gene_names):
    """utils.py:20-33"""
    if gene_names is None:
        if _is_anndata_like(data) and hasattr(data, "var_names"):
            gene_names = data.var_names
        elif isinstance(data, InputDataGLM):
            gene_names = data.features
        else:
            raise ValueError("Missing gene names")
    return np.asarray(gene_names)


def _shape_of(data):
    if isinstance(data, InputDataGLM):
        return data.x.shape
    if isinstance(data, CountShard):
        return (data.n_cells, data.n_genes)
    if _is_anndata_like(data):
        return data.X.shape
    return data.shape


def parse_sample_description(data, sample_description) -> pd.DataFrame:
    """utils.py:36-69"""
    if sample_description is None:
        if _is_anndata_like(data) and hasattr(data, "obs"):
            sample_description = data.obs
        else:
            raise ValueError(
                "Please specify `sample_description` or provide `data` as anndata.AnnData " +
                "with corresponding sample annotations"
            )
    assert _shape_of(data)[0] == sample_description.shape[0], \
        "data matrix and sample description must contain same number of cells: %i, %i" % \
        (_shape_of(data)[0], sample_description.shape[0])
    return sample_description


def parse_size_factors(size_factors, data, sample_description) -> Union[np.ndarray, None]:
    """utils.py:72-102"""
    if size_factors is not None:
        if isinstance(size_factors, pd.core.series.Series):
            size_factors = size_factors.values
        elif isinstance(size_factors, str):
            assert size_factors in sample_description.columns, ""
            size_factors = sample_description[size_factors].values
        size_factors = np.asarray(size_factors)
        assert size_factors.shape[0] == _shape_of(data)[0], "data matrix and size factors must contain same number of cells"
        assert np.all(size_factors > 0), "size_factors <= 0 found, please remove these cells"
    return size_factors


def parse_grouping(data, sample_description, grouping):
    """utils.py:105-109"""
    if isinstance(grouping, str):
        sample_description = parse_sample_description(data, sample_description)
        grouping = sample_description[grouping]
    return np.squeeze(np.asarray(grouping))


def split_x(data, grouping):
    """utils.py:112-117 -- groups in order of first appearance."""
    grouping = np.asarray(grouping)
    groups = pd.Series(grouping).unique()
    x0 = data[np.where(grouping == groups[0])[0]]
    x1 = data[np.where(grouping == groups[1])[0]]
    return x0, x1


def dmat_unique(dmat, sample_description):
    """utils.py:120-124"""
    dmat, idx = np.unique(dmat, axis=0, return_index=True)
    sample_description = sample_description.iloc[idx].reset_index(drop=True)
    return dmat, sample_description


def _as_categorical(sample_description, as_numeric):
    if isinstance(as_numeric, str):
        as_numeric = [as_numeric]
    if sample_description is not None:
        return [False if x in as_numeric else True for x in sample_description.columns.values]
    return True


def design_matrix(data=None, sample_description=None, formula=None, as_numeric=(), dmat=None, return_type="patsy"):
    """utils.py:127-176"""
    if data is None and sample_description is None and dmat is None:
        raise ValueError("supply either data or sample_description or dmat")
    if dmat is None and formula is None:
        raise ValueError("supply either dmat or formula")
    if dmat is None:
        sample_description = parse_sample_description(data, sample_description)
    return _design.design_matrix(sample_description=sample_description, formula=formula,
                                 as_categorical=_as_categorical(sample_description, as_numeric),
                                 dmat=dmat, return_type=return_type)


def preview_coef_names(sample_description, formula, as_numeric=()):
    """utils.py:179-212"""
    return _design.preview_coef_names(sample_description=sample_description, formula=formula,
                                      as_categorical=_as_categorical(sample_description, as_numeric))


def constraint_system_from_star(dmat=None, sample_description=None, formula=None, as_numeric=(),
                                constraints=None, return_type="patsy") -> Tuple:
    """utils.py:215-273"""
    return _design.constraint_system_from_star(dmat=dmat, sample_description=sample_description, formula=formula,
                                               as_categorical=_as_categorical(sample_description, as_numeric),
                                               constraints=constraints, return_type=return_type)


def bin_continuous_covariate(factor_to_bin, bins, data=None, sample_description=None):
    """utils.py:276-319"""
    if data is None and sample_description is not None:
        sd = sample_description
    elif data is not None and sample_description is None:
        sd = data.obs
    else:
        raise ValueError("supply either data or sample_description")
    if isinstance(bins, (list, np.ndarray, tuple)):
        bins = np.asarray(bins)
    else:
        bins = np.arange(0, 1, 1 / bins)
    fac_binned = _design.bin_continuous_covariate(sample_description=sd, factor_to_bin=factor_to_bin, bins=bins)
    sd[factor_to_bin + "_binned"] = fac_binned
    if data is None:
        return sample_description
