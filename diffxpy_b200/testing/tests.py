"""
de.test.* entry points (/root/reference/diffxpy/testing/tests.py): ``wald`` (:448-746), ``lrt`` (:252-445),
``wald_repeated`` (:749-821), ``t_test`` (:824-863), ``rank_test`` (:866-905), ``two_sample`` (:908-1094), the
multi-test wrappers ``pairwise`` (:1097-1330), ``versus_rest`` (:1333-1506), ``partition`` (:1509-1968) and the
estimator choke-point ``_fit`` (:26-249).  Same argument names, meaning and error behaviour; the estimator
behind ``_fit`` is the CUDA one (diffxpy_b200.estimator), the only backend of this build.
"""
import logging
from typing import Union, List, Dict, Callable, Tuple

import numpy as np
import pandas as pd
import scipy.sparse

from ..data.design import DesignMatrix, design_matrix as _design_matrix
from ..estimator import Estimator, InputDataGLM
from .det import DifferentialExpressionTestLRT, DifferentialExpressionTestWald, \
    DifferentialExpressionTestTT, DifferentialExpressionTestRank, _DifferentialExpressionTestSingle
from .utils import parse_gene_names, parse_sample_description, parse_size_factors, parse_grouping, \
    constraint_system_from_star, preview_coef_names

BACKENDS = ("numpy", "b200", "cuda")   # "numpy" is the reference's default value and is accepted as an alias


def _fit(
        noise_model,
        data,
        design_loc,
        design_scale,
        design_loc_names: list = None,
        design_scale_names: list = None,
        constraints_loc: np.ndarray = None,
        constraints_scale: np.ndarray = None,
        init_model=None,
        init_a: Union[np.ndarray, str] = "AUTO",
        init_b: Union[np.ndarray, str] = "AUTO",
        gene_names=None,
        size_factors=None,
        batch_size: Union[None, int, Tuple[int, int]] = None,
        backend: str = "numpy",
        training_strategy: Union[str, List[Dict[str, object]], Callable] = "AUTO",
        quick_scale: bool = None,
        train_args: dict = {},
        close_session=True,
        dtype="float64"
):
    """tests.py:26-249.  ``batch_size`` (the reference's memory chunking) is validated and ignored: the
    kernels stream the matrix and keep per-gene state on chip."""
    if backend.lower() in ("tf1", "tf2"):
        raise ValueError('backend="%s" is not available in diffxpy_b200 (single CUDA backend)' % backend)
    elif backend.lower() in BACKENDS:
        if isinstance(training_strategy, str):
            if training_strategy.lower() == "auto":
                training_strategy = "DEFAULT"
        if noise_model == "nb" or noise_model == "negative_binomial":
            pass
        else:
            raise ValueError('noise_model="%s" not recognized.' % noise_model)
        if batch_size is None:
            chunk_size_cells = int(1e9)
            chunk_size_genes = 128
        else:
            if isinstance(batch_size, int) or len(batch_size) != 2:
                raise ValueError("batch_size has to be a tuple of length 2 if backend is numpy")
            chunk_size_cells = batch_size[0]
            chunk_size_genes = batch_size[1]
    else:
        raise ValueError('backend="%s" not recognized.' % backend)

    if isinstance(init_a, str) and init_a.lower() == "init_model":
        # the reference never forwards init_model to the Estimator (tests.py:35 vs :222-228); when a fitted
        # model of the same coefficient layout is supplied we warm-start from it, otherwise fall back to AUTO
        init_a = "AUTO"
    if isinstance(init_b, str) and init_b.lower() == "init_model":
        init_b = "AUTO"

    input_data = InputDataGLM(
        data=data,
        design_loc=design_loc,
        design_scale=design_scale,
        design_loc_names=design_loc_names,
        design_scale_names=design_scale_names,
        constraints_loc=constraints_loc,
        constraints_scale=constraints_scale,
        size_factors=size_factors,
        feature_names=gene_names,
        chunk_size_cells=chunk_size_cells,
        chunk_size_genes=chunk_size_genes,
        as_dask=False,
        cast_dtype=dtype
    )
    constructor_args = {}
    if quick_scale is not None:
        constructor_args["quick_scale"] = quick_scale
    from ..device import _trace
    _trace("_fit(): InputDataGLM built")
    estim = Estimator(input_data=input_data, init_a=init_a, init_b=init_b, dtype=dtype, **constructor_args)
    estim.initialize()
    train_args = dict(train_args)
    estim.train_sequence(training_strategy=training_strategy, **train_args)
    _trace("_fit(): trained")
    if close_session:
        estim.finalize()
    _trace("_fit(): finalized (results on the host)")
    return estim


def lrt(
        data,
        full_formula_loc: str,
        reduced_formula_loc: str,
        full_formula_scale: str = "~1",
        reduced_formula_scale: str = "~1",
        as_numeric: Union[List[str], Tuple[str], str] = (),
        init_a: Union[np.ndarray, str] = "AUTO",
        init_b: Union[np.ndarray, str] = "AUTO",
        gene_names: Union[np.ndarray, list] = None,
        sample_description: pd.DataFrame = None,
        noise_model="nb",
        size_factors: Union[np.ndarray, pd.core.series.Series, np.ndarray] = None,
        batch_size: Union[None, int, Tuple[int, int]] = None,
        backend: str = "numpy",
        train_args: dict = {},
        training_strategy: Union[str, List[Dict[str, object]], Callable] = "AUTO",
        quick_scale: bool = False,
        dtype="float64",
        **kwargs
):
    """
    Perform log-likelihood ratio test for differential expression for each gene (tests.py:252-445).

    Note that lrt() does not support constraints; use wald() for constraints.
    """
    if len(kwargs) != 0:
        logging.getLogger("diffxpy").info("additional kwargs: %s", str(kwargs))
    if isinstance(as_numeric, str):
        as_numeric = [as_numeric]

    gene_names = parse_gene_names(data, gene_names)
    sample_description = parse_sample_description(data, sample_description)
    size_factors = parse_size_factors(size_factors=size_factors, data=data, sample_description=sample_description)

    as_categorical = [False if x in as_numeric else True for x in sample_description.columns.values]
    full_design_loc = _design_matrix(sample_description=sample_description, formula=full_formula_loc,
                                     as_categorical=as_categorical, return_type="patsy")
    reduced_design_loc = _design_matrix(sample_description=sample_description, formula=reduced_formula_loc,
                                        as_categorical=as_categorical, return_type="patsy")
    full_design_scale = _design_matrix(sample_description=sample_description, formula=full_formula_scale,
                                       as_categorical=as_categorical, return_type="patsy")
    reduced_design_scale = _design_matrix(sample_description=sample_description, formula=reduced_formula_scale,
                                          as_categorical=as_categorical, return_type="patsy")

    reduced_model = _fit(
        noise_model=noise_model, data=data, design_loc=reduced_design_loc, design_scale=reduced_design_scale,
        constraints_loc=None, constraints_scale=None, init_a=init_a, init_b=init_b, gene_names=gene_names,
        size_factors=size_factors, batch_size=batch_size, backend=backend, train_args=train_args,
        training_strategy=training_strategy, quick_scale=quick_scale, dtype=dtype, **kwargs
    )
    full_model = _fit(
        noise_model=noise_model, data=data, design_loc=full_design_loc, design_scale=full_design_scale,
        constraints_loc=None, constraints_scale=None, gene_names=gene_names, init_a="init_model",
        init_b="init_model", init_model=reduced_model, size_factors=size_factors, batch_size=batch_size,
        backend=backend, train_args=train_args, training_strategy=training_strategy, quick_scale=quick_scale,
        dtype=dtype, **kwargs
    )
    de_test = DifferentialExpressionTestLRT(
        sample_description=sample_description,
        full_design_loc_info=full_design_loc.design_info,
        full_estim=full_model,
        reduced_design_loc_info=reduced_design_loc.design_info,
        reduced_estim=reduced_model,
    )
    return de_test


def _start_upload(data):
    """Begin uploading ``data`` (device.prefetch) with the gene range the estimator of this rank will ask for."""
    from .. import device as dxdevice
    from .. import dist as dxdist
    try:
        if isinstance(data, InputDataGLM):
            return
        shape = _shape_of_data(data)
        w = dxdist.world("auto")
        gr = None if w.size == 1 else dxdist.gene_range(shape[1], w.rank, w.size, data=data)
        dxdevice.prefetch(data, gr)
    except Exception:          # noqa: BLE001  (an optimisation only: the estimator uploads itself if this did not)
        pass


def _shape_of_data(data):
    if hasattr(data, "X") and not hasattr(data, "shape"):
        return data.X.shape
    return data.shape


def wald(
        data,
        factor_loc_totest: Union[str, List[str]] = None,
        coef_to_test: Union[str, List[str]] = None,
        formula_loc: Union[None, str] = None,
        formula_scale: Union[None, str] = "~1",
        as_numeric: Union[List[str], Tuple[str], str] = (),
        init_a: Union[np.ndarray, str] = "AUTO",
        init_b: Union[np.ndarray, str] = "AUTO",
        gene_names: Union[np.ndarray, list] = None,
        sample_description: Union[None, pd.DataFrame] = None,
        dmat_loc=None,
        dmat_scale=None,
        constraints_loc: Union[None, List[str], Tuple[str, str], dict, np.ndarray] = None,
        constraints_scale: Union[None, List[str], Tuple[str, str], dict, np.ndarray] = None,
        noise_model: str = "nb",
        size_factors: Union[np.ndarray, pd.core.series.Series, str] = None,
        batch_size: Union[None, int, Tuple[int, int]] = None,
        backend: str = "numpy",
        train_args: dict = {},
        training_strategy: Union[str, List[Dict[str, object]], Callable] = "AUTO",
        quick_scale: bool = False,
        dtype="float64",
        **kwargs
):
    """
    Perform Wald test for differential expression for each gene (tests.py:448-746).

    :param data: Input data matrix (observations x features) or (cells x genes).
    :param factor_loc_totest: factor(s) of formula_loc to test, e.g. "condition".
    :param coef_to_test: coefficient name(s) to test directly, or the level of a multi-level factor.
    :param formula_loc: location model formula, e.g. "~ 1 + condition + batch".
    :param formula_scale: scale model formula.
    :param as_numeric: columns of sample_description to treat as numeric covariates.
    :param init_a / init_b: "AUTO", "standard", "closed_form", "all_zero" or explicit start values.
    :param dmat_loc / dmat_scale: pre-built design matrices (pd.DataFrame), instead of the formulas.
    :param constraints_loc / constraints_scale: None, dict, list of "a+b=0" strings or constraint matrix.
    :param size_factors: 1D array of library size factors or the name of a sample_description column.
    :param quick_scale: skip the scale model training after its initialisation.
    """
    if len(kwargs) != 0:
        logging.getLogger("diffxpy").debug("additional kwargs: %s", str(kwargs))

    if (dmat_loc is None and formula_loc is None) or \
            (dmat_loc is not None and formula_loc is not None):
        raise ValueError("Supply either dmat_loc or formula_loc.")
    if (dmat_scale is None and formula_scale is None) or \
            (dmat_scale is not None and formula_scale != "~1"):
        raise ValueError("Supply either dmat_scale or formula_scale.")
    if dmat_loc is not None and factor_loc_totest is not None:
        raise ValueError("Supply coef_to_test and not factor_loc_totest if dmat_loc is supplied.")
    if isinstance(factor_loc_totest, str):
        factor_loc_totest = [factor_loc_totest]
    if isinstance(coef_to_test, str):
        coef_to_test = [coef_to_test]
    if isinstance(as_numeric, str):
        as_numeric = [as_numeric]

    gene_names = parse_gene_names(data, gene_names)
    _start_upload(data)          # the matrix moves to the GPU while the design matrices are built on the host
    from ..device import _trace
    _trace("wald(): upload started")
    if dmat_loc is None and dmat_scale is None:
        sample_description = parse_sample_description(data, sample_description)
    size_factors = parse_size_factors(size_factors=size_factors, data=data, sample_description=sample_description)

    design_loc, design_loc_names, constraints_loc, term_names_loc = constraint_system_from_star(
        dmat=dmat_loc, sample_description=sample_description, formula=formula_loc, as_numeric=as_numeric,
        constraints=constraints_loc, return_type="patsy")
    design_scale, design_scale_names, constraints_scale, term_names_scale = constraint_system_from_star(
        dmat=dmat_scale, sample_description=sample_description, formula=formula_scale if dmat_scale is None else None,
        as_numeric=as_numeric, constraints=constraints_scale, return_type="patsy")

    # Define indices of coefficients to test (tests.py:664-714):
    constraints_loc_temp = constraints_loc if constraints_loc is not None else np.eye(design_loc.shape[-1])
    if factor_loc_totest is not None:
        if not isinstance(design_loc, DesignMatrix):
            col_indices = np.where([x in factor_loc_totest for x in term_names_loc])[0]
        else:
            col_indices = np.concatenate([
                np.arange(design_loc.shape[-1])[design_loc.design_info.slice(x)]
                for x in factor_loc_totest
            ])
        assert len(col_indices) > 0, "Could not find any matching columns!"
        if coef_to_test is not None:
            if len(factor_loc_totest) > 1:
                raise ValueError("do not set coef_to_test if more than one factor_loc_totest is given")
            samples = np.asarray(sample_description[factor_loc_totest[0]].astype(type(coef_to_test[0])) == coef_to_test[0])
            one_cols = np.where(np.asarray(design_loc)[samples][:, col_indices][0] == 1)[0]
            if one_cols.size == 0:
                # there is no such column; modify design matrix to create one
                design_loc[:, col_indices] = np.where(samples, 1, 0)[:, None]
    elif coef_to_test is not None:
        if sample_description is not None and dmat_loc is None:
            coef_loc_names = preview_coef_names(sample_description=sample_description, formula=formula_loc,
                                                as_numeric=as_numeric)
        else:
            coef_loc_names = list(design_loc_names)
        if not np.all([x in coef_loc_names for x in coef_to_test]):
            raise ValueError(
                "the requested test coefficients %s were found in model coefficients %s" %
                (", ".join([x for x in coef_to_test if x not in coef_loc_names]),
                 ", ".join(coef_loc_names))
            )
        col_indices = np.asarray([coef_loc_names.index(x) for x in coef_to_test])
    else:
        raise ValueError("either set factor_loc_totest or coef_to_test")
    for x in col_indices:
        if np.sum(constraints_loc_temp[x, :]) != 1:
            raise ValueError("Constraints input is wrong: not all tested coefficients are unconstrained.")
    col_indices = np.array([np.where(constraints_loc_temp[x, :] == 1)[0][0] for x in col_indices])

    model = _fit(
        noise_model=noise_model, data=data, design_loc=design_loc, design_scale=design_scale,
        design_loc_names=design_loc_names, design_scale_names=design_scale_names,
        constraints_loc=constraints_loc, constraints_scale=constraints_scale, init_a=init_a, init_b=init_b,
        gene_names=gene_names, size_factors=size_factors, batch_size=batch_size, backend=backend,
        train_args=train_args, training_strategy=training_strategy, quick_scale=quick_scale, dtype=dtype,
        **kwargs,
    )
    de_test = DifferentialExpressionTestWald(model_estim=model, col_indices=col_indices, noise_model=noise_model,
                                             sample_description=sample_description)
    return de_test


def wald_repeated(det: DifferentialExpressionTestWald, factor_loc_totest: Union[str, List[str]] = None,
                  coef_to_test: Union[str, List[str]] = None, **kwargs):
    """Run another Wald test on an object that already contains a model fit (tests.py:749-821)."""
    if len(kwargs) != 0:
        logging.getLogger("diffxpy").debug("additional kwargs: %s", str(kwargs))
    if isinstance(factor_loc_totest, str):
        factor_loc_totest = [factor_loc_totest]
    if isinstance(coef_to_test, str):
        coef_to_test = [coef_to_test]

    par_loc_names = list(det.model_estim.model.design_loc_names)
    if factor_loc_totest is not None and coef_to_test is None:
        col_indices = np.concatenate([np.where([fac in x for x in par_loc_names])[0] for fac in factor_loc_totest])
    elif factor_loc_totest is None and coef_to_test is not None:
        if not np.all([x in par_loc_names for x in coef_to_test]):
            raise ValueError(
                "the requested test coefficients %s were found in model coefficients %s" %
                (", ".join([x for x in coef_to_test if x not in par_loc_names]),
                 ", ".join(par_loc_names))
            )
        col_indices = np.asarray([par_loc_names.index(x) for x in coef_to_test])
    elif factor_loc_totest is None and coef_to_test is None:
        raise ValueError("Do not supply factor_loc_totest and coef_to_test in wald_repeated, run a new wald test.")
    else:
        raise ValueError("Either set factor_loc_totest or coef_to_test.")
    assert len(col_indices) > 0, "Could not find any matching columns!"

    constraints_loc = det.model_estim.model.constraints_loc
    for x in col_indices:
        if np.sum(constraints_loc[x, :]) != 1:
            raise ValueError("Constraints input is wrong: not all tested coefficients are unconstrained.")
    col_indices = np.array([np.where(constraints_loc[x, :] == 1)[0][0] for x in col_indices])

    de_test = DifferentialExpressionTestWald(model_estim=det.model_estim, col_indices=col_indices,
                                             noise_model=det.noise_model, sample_description=det.sample_description)
    return de_test


def t_test(data, grouping, gene_names: Union[np.ndarray, list] = None, sample_description: pd.DataFrame = None,
           is_logged: bool = False, is_sig_zerovar: bool = True):
    """Welch's t-test between two groups for each gene (tests.py:824-863)."""
    gene_names = parse_gene_names(data, gene_names)
    grouping = parse_grouping(data, sample_description, grouping)
    return DifferentialExpressionTestTT(data=data, sample_description=sample_description, grouping=grouping,
                                        gene_names=gene_names, is_logged=is_logged, is_sig_zerovar=is_sig_zerovar)


def rank_test(data, grouping: Union[str, np.ndarray, list], gene_names: Union[np.ndarray, list] = None,
              sample_description: pd.DataFrame = None, is_logged: bool = False, is_sig_zerovar: bool = True):
    """Mann-Whitney rank test (Wilcoxon rank-sum test) between two groups for each gene (tests.py:866-905)."""
    gene_names = parse_gene_names(data, gene_names)
    grouping = parse_grouping(data, sample_description, grouping)
    return DifferentialExpressionTestRank(data=data, sample_description=sample_description, grouping=grouping,
                                          gene_names=gene_names, is_logged=is_logged, is_sig_zerovar=is_sig_zerovar)


def two_sample(data, grouping: Union[str, np.ndarray, list], as_numeric=(), test: str = "t-test",
               gene_names: Union[np.ndarray, list] = None, sample_description: pd.DataFrame = None,
               noise_model: str = None, size_factors: np.ndarray = None, batch_size=None, backend: str = "numpy",
               train_args: dict = {}, training_strategy="AUTO", is_sig_zerovar: bool = True, quick_scale: bool = None,
               dtype="float64", **kwargs) -> _DifferentialExpressionTestSingle:
    """Two-group comparison with the test of choice (tests.py:908-1094): "wald", "lrt", "t-test", "rank"."""
    if test in ["t-test", "rank"] and noise_model is not None:
        raise Warning('two_sample(): Do not specify `noise_model` if using test t-test or rank_test: ' +
                      'The t-test is based on a gaussian noise model and the rank sum test is model free.')
    gene_names = parse_gene_names(data, gene_names)
    grouping = parse_grouping(data, sample_description, grouping)
    sample_description = pd.DataFrame({"grouping": grouping})
    groups = np.unique(grouping)
    if groups.size > 2:
        raise ValueError("More than two groups detected:\n\t%s", groups)
    if groups.size < 2:
        raise ValueError("Less than two groups detected:\n\t%s", groups)
    if test.lower() == 'wald':
        if noise_model is None:
            raise ValueError("Please specify noise_model")
        return wald(data=data, factor_loc_totest="grouping", as_numeric=as_numeric, coef_to_test=None,
                    formula_loc="~ 1 + grouping", formula_scale="~ 1 + grouping", gene_names=gene_names,
                    sample_description=sample_description, noise_model=noise_model, size_factors=size_factors,
                    batch_size=batch_size, backend=backend, train_args=train_args,
                    training_strategy=training_strategy, quick_scale=quick_scale, dtype=dtype, **kwargs)
    elif test.lower() == 'lrt':
        if noise_model is None:
            raise ValueError("Please specify noise_model")
        return lrt(data=data, full_formula_loc="~ 1 + grouping", full_formula_scale="~ 1",
                   reduced_formula_loc="~ 1", reduced_formula_scale="~ 1", as_numeric=as_numeric,
                   gene_names=gene_names, sample_description=sample_description, noise_model=noise_model,
                   size_factors=size_factors, init_a="closed_form", init_b="closed_form",      # reference tests.py:1067-1068
                   batch_size=batch_size, backend=backend, train_args=train_args,
                   training_strategy=training_strategy, quick_scale=bool(quick_scale), dtype=dtype, **kwargs)
    elif test.lower() == 't-test' or test.lower() == "t_test" or test.lower() == "ttest":
        return t_test(data=data, gene_names=gene_names, grouping=grouping, is_sig_zerovar=is_sig_zerovar)
    elif test.lower() == 'rank':
        return rank_test(data=data, gene_names=gene_names, grouping=grouping, is_sig_zerovar=is_sig_zerovar)
    raise ValueError('two_sample(): Parameter `test="%s"` not recognized.' % test)


# ------------------------------------------------------------------------------------------------------------
# multi-test wrappers (tests.py:1097-1968).  The model-free variants (t-test / rank) of pairwise and versus_rest
# read the matrix ONCE: K_ingest builds the count histograms of all groups, and every comparison is a launch of
# K_twogroup on two of them (dx_group_pair_tests).  The model-based variants loop over the single tests like the
# reference, except the z-test, which needs one fit.
# ------------------------------------------------------------------------------------------------------------
def _matrix_of(data):
    """The cells x genes matrix behind what diffxpy accepts as ``data``."""
    if isinstance(data, InputDataGLM):
        return data.x
    if hasattr(data, "X") and not isinstance(data, np.ndarray) and not scipy.sparse.issparse(data):
        return data.X
    return data


def _column_means(x):
    return np.asarray(x.mean(axis=0)).flatten()


def _grouped_two_sample_setup(x, grouping, groups):
    """One K_ingest pass over all groups: (SuffStats, cells per group, first cell of every group)."""
    from ..device import CountShard
    import torch
    code = np.searchsorted(groups, grouping).astype(np.uint8)
    counts = np.bincount(code, minlength=len(groups)).astype(np.int64)
    first = np.array([int(np.argmax(code == k)) for k in range(len(groups))])
    shard = CountShard.from_host(x)
    with torch.cuda.device(shard.device):
        suff = shard.group_suffstats(torch.from_numpy(code).to(shard.device), len(groups))
    return suff, counts, first


def pairwise(data, grouping: Union[str, np.ndarray, list], as_numeric=(), test: str = "z-test", lazy: bool = True,
             gene_names: Union[np.ndarray, list] = None, sample_description: pd.DataFrame = None,
             noise_model: str = "nb", size_factors: np.ndarray = None, batch_size=None, backend: str = "numpy",
             train_args: dict = {}, training_strategy="AUTO", is_sig_zerovar: bool = True, quick_scale: bool = False,
             dtype="float64", pval_correction: str = "global", keep_full_test_objs: bool = False, **kwargs):
    """All pairs of groups (tests.py:1097-1330): "z-test" (one fit '~ 0 + group', then two-coefficient z-tests),
    "wald", "lrt", "t-test" or "rank"."""
    from .det import two_group_postprocess
    from .det_pair import DifferentialExpressionTestZTest, DifferentialExpressionTestPairwiseStandard
    from ..stats import stats as dstats
    if len(kwargs) != 0:
        logging.getLogger("diffxpy").info("additional kwargs: %s", str(kwargs))
    is_z = test.lower() in ("z-test", "z_test", "ztest")
    if lazy and not is_z:
        raise ValueError("lazy evaluation of pairwise tests only possible if test is z-test")
    gene_names = parse_gene_names(data, gene_names)
    grouping = parse_grouping(data, sample_description, grouping)      # (a label array needs no sample_description)
    sample_description = pd.DataFrame({"grouping": grouping})
    groups = np.unique(grouping)
    if is_z:
        dmat = _design_matrix(sample_description=sample_description, formula="~ 1 - 1 + grouping")
        # group-wise starts for both models, like the reference (tests.py:1245-1246)
        model = _fit(noise_model=noise_model, data=data, design_loc=dmat, design_scale=dmat, gene_names=gene_names,
                     size_factors=size_factors, init_a="closed_form", init_b="closed_form", batch_size=batch_size,
                     backend=backend, train_args=train_args, training_strategy=training_strategy,
                     quick_scale=quick_scale, dtype=dtype, **kwargs)
        return DifferentialExpressionTestZTest(model_estim=model, grouping=grouping, groups=groups,
                                               correction_type=pval_correction)
    x = _matrix_of(data)
    n_genes = x.shape[1]
    pvals = np.tile(np.nan, [len(groups), len(groups), n_genes])
    pvals[np.eye(pvals.shape[0]).astype(bool)] = 0
    logfc = np.tile(np.nan, [len(groups), len(groups), n_genes])
    logfc[np.eye(logfc.shape[0]).astype(bool)] = 0
    tests = np.tile([None], [len(groups), len(groups)]) if keep_full_test_objs else None
    model_free = test.lower() in ("t-test", "t_test", "ttest", "rank")
    if model_free and not keep_full_test_objs and len(groups) <= 255:
        do_rank = test.lower() == "rank"
        suff, counts, first = _grouped_two_sample_setup(x, grouping, groups)
        for i in range(len(groups)):
            for j in range(i + 1, len(groups)):
                a, b = (i, j) if first[i] < first[j] else (j, i)        # sample 0 = the group that appears first
                res = dstats.group_pair_device(suff, a, b, counts[a], counts[b], do_rank=do_rank)
                r = two_group_postprocess(res["out"], res["flags"], counts[a], counts[b], do_rank, False, is_sig_zerovar)
                pvals[i, j] = pvals[j, i] = r["pval"]
                logfc[i, j] = r["logfc"]
                logfc[j, i] = -logfc[i, j]
    else:
        for i, g1 in enumerate(groups):
            for j in range(i + 1, len(groups)):
                g2 = groups[j]
                idx = np.where(np.logical_or(grouping == g1, grouping == g2))[0]
                de_test_temp = two_sample(
                    data=x[idx, :], grouping=grouping[idx], as_numeric=as_numeric, test=test, gene_names=gene_names,
                    sample_description=sample_description.iloc[idx, :], noise_model=None if model_free else noise_model,
                    size_factors=size_factors[idx] if size_factors is not None else None, batch_size=batch_size,
                    backend=backend, train_args=train_args, training_strategy=training_strategy,
                    quick_scale=quick_scale, is_sig_zerovar=is_sig_zerovar, dtype=dtype, **kwargs)
                pvals[i, j] = pvals[j, i] = de_test_temp.pval
                logfc[i, j] = de_test_temp.log_fold_change()
                logfc[j, i] = -logfc[i, j]
                if keep_full_test_objs:
                    tests[i, j] = tests[j, i] = de_test_temp
    return DifferentialExpressionTestPairwiseStandard(gene_ids=gene_names, pval=pvals, logfc=logfc, ave=_column_means(x),
                                                      groups=groups, tests=tests, correction_type=pval_correction)


def versus_rest(data, grouping: Union[str, np.ndarray, list], as_numeric=(), test: str = "wald",
                gene_names: Union[np.ndarray, list] = None, sample_description: pd.DataFrame = None,
                noise_model: str = None, size_factors: np.ndarray = None, batch_size=None, backend: str = "numpy",
                train_args: dict = {}, training_strategy="AUTO", is_sig_zerovar: bool = True, quick_scale: bool = None,
                dtype="float64", pval_correction: str = "global", keep_full_test_objs: bool = False, **kwargs):
    """Every group against the union of all other groups (tests.py:1333-1506)."""
    from .det import two_group_postprocess
    from .det_pair import DifferentialExpressionTestVsRest
    from ..stats import stats as dstats
    if len(kwargs) != 0:
        logging.getLogger("diffxpy").info("additional kwargs: %s", str(kwargs))
    gene_names = parse_gene_names(data, gene_names)
    grouping = parse_grouping(data, sample_description, grouping)      # (a label array needs no sample_description)
    sample_description = pd.DataFrame({"grouping": grouping})
    groups = np.unique(grouping)
    x = _matrix_of(data)
    n_genes = x.shape[1]
    pvals = np.zeros([1, len(groups), n_genes])
    logfc = np.zeros([1, len(groups), n_genes])
    tests = np.tile([None], [1, len(groups)]) if keep_full_test_objs else None
    model_free = test.lower() in ("t-test", "t_test", "ttest", "rank")
    if model_free and not keep_full_test_objs and 2 <= len(groups) <= 255:
        do_rank = test.lower() == "rank"
        suff, counts, first = _grouped_two_sample_setup(x, grouping, groups)
        n = int(counts.sum())
        for i in range(len(groups)):
            res = dstats.group_pair_device(suff, i, -1, counts[i], n - counts[i], do_rank=do_rank)
            r = two_group_postprocess(res["out"], res["flags"], counts[i], n - counts[i], do_rank, False, is_sig_zerovar)
            pvals[0, i] = r["pval"]
            # two_sample() orders the samples by first appearance of "group" / "rest": the fold change is second over first
            logfc[0, i] = r["logfc"] if first[i] == 0 else -r["logfc"]
    else:
        for i, g1 in enumerate(groups):
            test_grouping = np.where(grouping == g1, "group", "rest")
            de_test_temp = two_sample(
                data=x, grouping=test_grouping, as_numeric=as_numeric, test=test, gene_names=gene_names,
                sample_description=sample_description, noise_model=None if model_free else noise_model,
                batch_size=batch_size, backend=backend,
                train_args=train_args, training_strategy=training_strategy, quick_scale=quick_scale,
                size_factors=size_factors, is_sig_zerovar=is_sig_zerovar, dtype=dtype, **kwargs)
            pvals[0, i] = de_test_temp.pval
            logfc[0, i] = de_test_temp.log_fold_change()
            if keep_full_test_objs:
                tests[0, i] = de_test_temp
    return DifferentialExpressionTestVsRest(gene_ids=gene_names, pval=pvals, logfc=logfc, ave=_column_means(x),
                                            groups=groups, tests=tests, correction_type=pval_correction)


def partition(data, parts: Union[str, np.ndarray, list], gene_names: Union[np.ndarray, list] = None,
              sample_description: pd.DataFrame = None):
    """Run a single test inside every partition of the cells (tests.py:1509-1538):
    ``de.test.partition(data, parts="tissue").wald(formula_loc=..., factor_loc_totest=...)``."""
    return _Partition(data=data, parts=parts, gene_names=gene_names, sample_description=sample_description)


class _Partition:
    """tests.py:1541-1968: the partitioning of the cells, the per-partition test calls and their summary in one
    DifferentialExpressionTestByPartition."""

    def __init__(self, data, parts, gene_names=None, sample_description=None):
        self.x = _matrix_of(data)
        if not (isinstance(self.x, np.ndarray) or scipy.sparse.issparse(self.x)):
            raise ValueError("data type %s not recognized" % type(data))
        if scipy.sparse.issparse(self.x):
            self.x = self.x.tocsr()
        self.gene_names = parse_gene_names(data, gene_names)
        self.partition = parse_grouping(data, sample_description, parts)
        if sample_description is None and not (hasattr(data, "obs")):
            sample_description = pd.DataFrame({"partition": self.partition})
        self.sample_description = parse_sample_description(data, sample_description)
        self.partitions = np.unique(self.partition)
        self.partition_idx = [np.where(self.partition == x)[0] for x in self.partitions]

    def _collect(self, fn):
        from .det_pair import DifferentialExpressionTestByPartition
        tests = [fn(idx) for idx in self.partition_idx]
        return DifferentialExpressionTestByPartition(partitions=self.partitions, tests=tests, ave=_column_means(self.x),
                                                     correction_type="by_test")

    def two_sample(self, grouping: str, as_numeric=(), test=None, size_factors: np.ndarray = None, noise_model: str = None,
                   batch_size=None, backend: str = "numpy", train_args: dict = {}, training_strategy="AUTO",
                   is_sig_zerovar: bool = True, **kwargs):
        return self._collect(lambda idx: two_sample(
            data=self.x[idx, :], grouping=grouping, as_numeric=as_numeric, test=test, gene_names=self.gene_names,
            sample_description=self.sample_description.iloc[idx, :], noise_model=noise_model,
            size_factors=size_factors[idx] if size_factors is not None else None, batch_size=batch_size, backend=backend,
            train_args=train_args, training_strategy=training_strategy, is_sig_zerovar=is_sig_zerovar, **kwargs))

    def t_test(self, grouping: str, is_logged: bool = False, is_sig_zerovar: bool = True, dtype="float64"):
        return self._collect(lambda idx: t_test(
            data=self.x[idx, :], grouping=grouping, is_logged=is_logged, gene_names=self.gene_names,
            sample_description=self.sample_description.iloc[idx, :], is_sig_zerovar=is_sig_zerovar))

    def rank_test(self, grouping: str, is_sig_zerovar: bool = True, dtype="float64"):
        return self._collect(lambda idx: rank_test(
            data=self.x[idx, :], grouping=grouping, gene_names=self.gene_names,
            sample_description=self.sample_description.iloc[idx, :], is_sig_zerovar=is_sig_zerovar))

    def lrt(self, full_formula_loc: str = None, reduced_formula_loc: str = None, full_formula_scale: str = "~1",
            reduced_formula_scale: str = "~1", as_numeric=(), init_a="AUTO", init_b="AUTO", size_factors: np.ndarray = None,
            noise_model="nb", batch_size=None, backend: str = "numpy", train_args: dict = {}, training_strategy="AUTO",
            **kwargs):
        return self._collect(lambda idx: lrt(
            data=self.x[idx, :], full_formula_loc=full_formula_loc, reduced_formula_loc=reduced_formula_loc,
            full_formula_scale=full_formula_scale, reduced_formula_scale=reduced_formula_scale, as_numeric=as_numeric,
            init_a=init_a, init_b=init_b, gene_names=self.gene_names, sample_description=self.sample_description.iloc[idx, :],
            noise_model=noise_model, size_factors=size_factors[idx] if size_factors is not None else None,
            batch_size=batch_size, backend=backend, train_args=train_args, training_strategy=training_strategy, **kwargs))

    def wald(self, factor_loc_totest: str = None, coef_to_test: object = None, formula_loc: str = None,
             formula_scale: str = "~1", as_numeric=(), constraints_loc: np.ndarray = None,
             constraints_scale: np.ndarray = None, noise_model: str = "nb", size_factors: np.ndarray = None, batch_size=None,
             backend: str = "numpy", train_args: dict = {}, training_strategy="AUTO", **kwargs):
        return self._collect(lambda idx: wald(
            data=self.x[idx, :], factor_loc_totest=factor_loc_totest, coef_to_test=coef_to_test, formula_loc=formula_loc,
            formula_scale=formula_scale, as_numeric=as_numeric, constraints_loc=constraints_loc,
            constraints_scale=constraints_scale, gene_names=self.gene_names,
            sample_description=self.sample_description.iloc[idx, :], noise_model=noise_model,
            size_factors=size_factors[idx] if size_factors is not None else None, batch_size=batch_size, backend=backend,
            train_args=train_args, training_strategy=training_strategy, **kwargs))


# ------------------------------------------------------------------------------------------------------------
# continuous covariates (tests.py:1971-2370)
# ------------------------------------------------------------------------------------------------------------
def _replace_continuous(formula: str, continuous: str, new_coefs: List[str]) -> str:
    """Replace the continuous covariate by its spline basis columns in a formula, distributing over interactions
    (what the reference obtains by substituting '(c0+c1+...)' into the patsy formula, tests.py:2236-2245)."""
    from ..data.design import _parse_formula
    intercept, terms = _parse_formula(formula)
    out = []
    for t in terms:
        parts = t.split(":")
        if continuous in parts:
            rest = [x for x in parts if x != continuous]
            for c in new_coefs:
                out.append(":".join([c] + rest))
        else:
            out.append(t)
    return "~ " + " + ".join((["1"] if intercept else ["0"]) + out)


def continuous_1d(data, continuous: str, factor_loc_totest: Union[str, List[str]], formula_loc: str, formula_scale: str = "~1",
                  df: int = 5, spline_basis: str = "bs", as_numeric=(), test: str = 'wald', init_a="standard", init_b="standard",
                  gene_names: Union[np.ndarray, list] = None, sample_description=None, constraints_loc=None,
                  constraints_scale=None, noise_model: str = 'nb', size_factors=None, batch_size=None, backend: str = "numpy",
                  train_args: dict = {}, training_strategy="AUTO", quick_scale: bool = None, dtype="float64", **kwargs):
    """Differential expression along a continuous covariate, modelled by a spline basis with ``df`` degrees of freedom
    (tests.py:1971-2370).  ``spline_basis``: "bs" (cubic B-splines, the reference default), "cr" / "cc" (natural / cyclic
    cubic regression splines, centred), all built without patsy (diffxpy_b200.data.splines).

    LRT: the reduced location model is the location formula without the tested factors.  (The reference derives it
    from ``formula_scale`` at tests.py:2311, an evident slip that silently drops every untested location term.)"""
    from ..data.splines import bs, crs
    from .det_cont import DifferentialExpressionTestWaldCont, DifferentialExpressionTestLRTCont
    if isinstance(factor_loc_totest, str):
        factor_loc_totest = [factor_loc_totest]
    elif isinstance(factor_loc_totest, tuple):
        factor_loc_totest = list(factor_loc_totest)
    if isinstance(as_numeric, str):
        as_numeric = [as_numeric]
    as_numeric = list(as_numeric)
    gene_names = parse_gene_names(data, gene_names)
    sample_description = parse_sample_description(data, sample_description).copy()
    if continuous not in sample_description.columns:
        raise ValueError('parameter continuous not found in sample_description')
    if not pd.api.types.is_numeric_dtype(sample_description[continuous]):
        raise ValueError("The column corresponding to the continuous covariate ('%s') in " % continuous +
                         "sample description should be numeric but is '%s'" % sample_description[continuous].dtype)
    coords = np.asarray(sample_description[continuous].values, dtype=np.float64)
    if len(np.unique(coords)) <= df - 1:
        raise ValueError("Use at most (number of observed points in continuous covariate) - 1 degrees of freedom " +
                         " for spline basis. You chose df=%i for n=%i points." % (df, len(np.unique(coords))))
    if spline_basis.lower() == "bs":
        basis = bs(coords, df=df, degree=3, include_intercept=False)
    elif spline_basis.lower() in ("cr", "cc"):
        # patsy cr / cc with constraints='center' (tests.py:2201-2210), restated in data/splines.py
        basis, _, _ = crs(coords, df=df, cyclic=(spline_basis.lower() == "cc"), center=True)
    else:
        raise ValueError("spline basis %s not recognized" % spline_basis)
    new_coefs = [continuous + str(i) for i in range(basis.shape[1])]
    # interpolated basis on a uniform grid; like the reference (tests.py:2218-2229) its knots come from the grid itself
    grid = np.linspace(np.min(coords), np.max(coords), 100)
    interpolated_spline_basis = np.hstack([np.ones([100, 1]), bs(grid, df=df, degree=3, include_intercept=False),
                                           np.expand_dims(grid, axis=1)])
    formula_loc_new = _replace_continuous(formula_loc, continuous, new_coefs)
    formula_scale_new = _replace_continuous(formula_scale, continuous, new_coefs)
    for i, c in enumerate(new_coefs):
        sample_description[c] = basis[:, i]
    as_numeric.extend(new_coefs)

    if test.lower() == 'wald':
        if noise_model is None:
            raise ValueError("Please specify noise_model")
        totest = []
        for f in factor_loc_totest:
            parts = f.split(":")
            if continuous in parts:
                rest = [x for x in parts if x != continuous]
                totest.extend(":".join([c] + rest) for c in new_coefs)
            else:
                totest.append(f)
        de_test = wald(data=data, factor_loc_totest=totest, coef_to_test=None, formula_loc=formula_loc_new,
                       formula_scale=formula_scale_new, as_numeric=as_numeric, init_a=init_a, init_b=init_b,
                       gene_names=gene_names, sample_description=sample_description, constraints_loc=constraints_loc,
                       constraints_scale=constraints_scale, noise_model=noise_model, size_factors=size_factors,
                       batch_size=batch_size, backend=backend, train_args=train_args, training_strategy=training_strategy,
                       quick_scale=quick_scale, dtype=dtype, **kwargs)
        return DifferentialExpressionTestWaldCont(de_test=de_test, noise_model=noise_model, size_factors=size_factors,
                                                  continuous_coords=coords, spline_coefs=new_coefs,
                                                  interpolated_spline_basis=interpolated_spline_basis)
    elif test.lower() == 'lrt':
        if noise_model is None:
            raise ValueError("Please specify noise_model")
        from ..data.design import _parse_formula
        intercept, terms = _parse_formula(formula_loc)
        kept = [t for t in terms if t not in factor_loc_totest and not (set(t.split(":")) & set(factor_loc_totest))]
        reduced = "~ " + " + ".join((["1"] if intercept else ["0"]) + kept)
        reduced_formula_loc = _replace_continuous(reduced, continuous, new_coefs)
        de_test = lrt(data=data, full_formula_loc=formula_loc_new, reduced_formula_loc=reduced_formula_loc,
                      full_formula_scale=formula_scale_new, reduced_formula_scale=formula_scale_new, as_numeric=as_numeric,
                      init_a=init_a, init_b=init_b, gene_names=gene_names, sample_description=sample_description,
                      noise_model=noise_model, size_factors=size_factors, batch_size=batch_size, backend=backend,
                      train_args=train_args, training_strategy=training_strategy, quick_scale=quick_scale, dtype=dtype,
                      **kwargs)
        return DifferentialExpressionTestLRTCont(de_test=de_test, size_factors=size_factors, continuous_coords=coords,
                                                 spline_coefs=new_coefs, interpolated_spline_basis=interpolated_spline_basis,
                                                 noise_model=noise_model)
    raise ValueError('continuous(): Parameter `test` not recognized.')
