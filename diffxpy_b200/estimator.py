"""
The batchglm Estimator protocol, served by the CUDA kernels of libdxb200.so.

This is the drop-in boundary: diffxpy's ``_fit``
(/root/reference/diffxpy/testing/tests.py:26-249) builds ``InputDataGLM(...)``, then
``Estimator(input_data, init_a, init_b, dtype, quick_scale)``, calls ``initialize()``,
``train_sequence(training_strategy, **train_args)``, ``finalize()`` and afterwards reads ``a_var``,
``fisher_inv``, ``log_likelihood``, ``jacobian``, ``error_codes``, ``niter``, ``x``, ``input_data.*`` and
``model.*`` (det.py:489-513, 714-765, 792-811).  The classes below keep exactly that surface.
"""
import logging

import numpy as np
import pandas as pd
import scipy.sparse
import torch

from . import _lib
from .device import CountShard, ptr, stream_ptr, cuda_device
from . import device as dxdevice
from . import dist as dxdist

logger = logging.getLogger("diffxpy")

TRAINING_STRATEGIES = {
    # batchglm numpy TrainingStrategies.DEFAULT; method_b / ftol_b / max_iter_b belong to the reference's
    # scipy brent scale step and are accepted but unused: the scale step here is a safeguarded Newton.
    "DEFAULT": [{"max_steps": 1000, "method_b": "brent", "update_b_freq": 5, "ftol_b": 1e-6, "max_iter_b": 1000}],
    "QUICK": [{"max_steps": 100, "method_b": "brent", "update_b_freq": 5, "ftol_b": 1e-4, "max_iter_b": 1000}],
}
LLTOL_BY_FEATURE = 1e-10
_B_CLOSED_FORM = -1      # host-side marker: init_b="closed_form", turned into DX_INIT_GIVEN before the first launch


def _dense_design(d):
    if hasattr(d, "values") and not isinstance(d, np.ndarray):
        d = d.values
    return np.ascontiguousarray(np.asarray(d, dtype=np.float64))


def _names_of(d, names, prefix):
    if names is not None:
        return list(names)
    di = getattr(d, "design_info", None)
    if di is not None:
        return list(di.column_names)
    if hasattr(d, "columns"):
        return [str(c) for c in d.columns]
    return ["%s%d" % (prefix, i) for i in range(np.asarray(d).shape[1])]


_HASH_MULT = np.array([0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0xD6E8FEB86659FD93,
                       0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53, 0x2545F4914F6CDD1D, 0x9FB21C651E98DF25], dtype=np.uint64)


def _group_rows_exact(mat):
    v = mat.view(np.dtype((np.void, mat.dtype.itemsize * mat.shape[1]))).ravel()
    _, idx, inv = np.unique(v, return_index=True, return_inverse=True)
    return mat[idx], np.asarray(inv).reshape(-1)


def group_rows(mat):
    """Distinct rows of a float64 matrix and the row -> group map (byte-wise; -0.0 normalised).  Groups are
    numbered in the byte order of the rows, like ``np.unique`` on the raw records.  Rows are reduced to a 64-bit
    multiplicative hash (integer matrix-vector product: no sort of wide records, no 2-d temporary) and grouped with a hash
    table; a second, independent 64-bit hash must be constant inside every group (a wrong merge would need a simultaneous
    collision of both, ~n^2 2^-128), else -- and for anything unusual -- the exact method takes over."""
    mat = np.ascontiguousarray(mat + 0.0)
    if mat.shape[1] == 0:
        return mat[:1], np.zeros(mat.shape[0], dtype=np.int64)
    if mat.shape[0] < 4096:
        return _group_rows_exact(mat)
    bits = mat.view(np.uint64)
    # (0.0 / 1.0 / 2.0 ... differ in their top bits only, which a multiplication pushes out of the word: fold the upper
    # half onto the lower one first -- a bijection per element, so equal rows stay equal and distinct rows distinct)
    mixed = bits ^ (bits >> np.uint64(32))
    odd = 2 * np.arange(mat.shape[1], dtype=np.uint64) + np.uint64(1)
    with np.errstate(over="ignore"):
        h = mixed @ (np.resize(_HASH_MULT, mat.shape[1]) * odd)
        h ^= h >> np.uint64(29)
        h2 = mixed @ (np.resize(_HASH_MULT[::-1], mat.shape[1]) * (odd + np.uint64(2 * mat.shape[1])))
    # hash table instead of a sort: codes in order of first appearance, idx = first row of every distinct hash
    inv, uniq_h = pd.factorize(h)
    inv = np.asarray(inv, dtype=np.int64)
    if uniq_h.shape[0] > 65536:
        return _group_rows_exact(mat)
    idx = np.empty(uniq_h.shape[0], dtype=np.int64)
    idx[inv[::-1]] = np.arange(mat.shape[0] - 1, -1, -1)          # (the last write wins: the first row of every code)
    if not np.array_equal(h2[idx][inv], h2):
        return _group_rows_exact(mat)
    reps = mat[idx]
    # renumber in the byte order of the representative rows (what the exact method returns)
    uniq, order_inv = _group_rows_exact(reps)
    if uniq.shape[0] != reps.shape[0]:
        return _group_rows_exact(mat)
    return uniq, order_inv[inv]


def join_groupings(parts):
    """Grouping of the rows of ``np.concatenate([m_0, m_1, ...], axis=1)`` from the groupings ``(uniq_k, inv_k)`` of its column
    blocks -- the same (uniq, inv) as group_rows() on the concatenation (groups in the byte order of the joint rows: every
    uniq_k is in byte order, so the joint order is the lexicographic order of the index tuples), without hashing the wide
    rows again.  None when the number of index tuples is too large for a counting pass."""
    span = 1
    for uniq_k, _ in parts:
        span *= int(uniq_k.shape[0])
    if span > (1 << 22):
        return None
    code = np.zeros(parts[0][1].shape[0], dtype=np.int64)
    for uniq_k, inv_k in parts:
        code *= int(uniq_k.shape[0])
        code += inv_k
    present = np.flatnonzero(np.bincount(code, minlength=span))          # ascending codes = byte order of the joint rows
    remap = np.zeros(span, dtype=np.int64)
    remap[present] = np.arange(present.shape[0])
    cols, rest = [], present.copy()
    for uniq_k, _ in reversed(parts):
        cols.append(uniq_k[rest % int(uniq_k.shape[0])])
        rest //= int(uniq_k.shape[0])
    return np.ascontiguousarray(np.concatenate(cols[::-1], axis=1)), remap[code]


def few_distinct_rows(mat, limit=4096, probe=8192):
    """group_rows(mat) when the matrix has at most ``limit`` distinct rows, else None (decided on a prefix first, so a
    continuous design costs one small probe instead of a sort of all its rows)."""
    if mat.shape[0] > probe and group_rows(mat[:probe])[0].shape[0] > limit:
        return None
    g = group_rows(mat)
    return g if g[0].shape[0] <= limit else None


def rank_of_rows(mat, groups=None):
    """Rank of a tall matrix through its distinct rows (same row space) when they are few."""
    if groups is not None and groups[0].shape[0] <= 4096:
        return int(np.linalg.matrix_rank(groups[0]))
    return int(np.linalg.matrix_rank(mat))


class InputDataGLM:
    """batchglm ``InputDataGLM`` (constructed at tests.py:177-191)."""

    def __init__(self, data, design_loc, design_scale, design_loc_names=None, design_scale_names=None,
                 constraints_loc=None, constraints_scale=None, size_factors=None, feature_names=None,
                 chunk_size_cells=int(1e9), chunk_size_genes=128, as_dask=True, cast_dtype="float64"):
        if hasattr(data, "X") and not isinstance(data, np.ndarray) and not scipy.sparse.issparse(data) \
                and not isinstance(data, CountShard):
            if feature_names is None and hasattr(data, "var_names"):
                feature_names = np.asarray(data.var_names)
            data = data.X
        self.data = data
        self.x = data
        self.loc_names = _names_of(design_loc, design_loc_names, "loc")
        self.scale_names = _names_of(design_scale, design_scale_names, "scale")
        self.design_loc_names = self.loc_names
        self.design_scale_names = self.scale_names
        self.design_loc = _dense_design(design_loc)
        self.design_scale = _dense_design(design_scale)
        n = self.design_loc.shape[0]
        shape = (data.n_cells, data.n_genes) if isinstance(data, CountShard) else data.shape
        if shape[0] != n or self.design_scale.shape[0] != n:
            raise ValueError("data and design matrices must contain the same number of observations")
        self.constraints_loc = np.eye(self.design_loc.shape[1]) if constraints_loc is None \
            else np.asarray(constraints_loc, dtype=np.float64)
        self.constraints_scale = np.eye(self.design_scale.shape[1]) if constraints_scale is None \
            else np.asarray(constraints_scale, dtype=np.float64)
        xh_loc = self.design_loc if constraints_loc is None else self.design_loc @ self.constraints_loc
        xh_scale = self.design_scale if constraints_scale is None else self.design_scale @ self.constraints_scale
        # distinct rows of the constrained designs: rank check here, design groups in Estimator.initialize()
        self._groups_loc = few_distinct_rows(xh_loc) if n >= 4096 else None
        self._groups_scale = few_distinct_rows(xh_scale) if n >= 4096 else None
        r_loc = rank_of_rows(xh_loc, self._groups_loc)
        if r_loc != xh_loc.shape[1]:
            raise ValueError("constrained design matrix design_loc is not full rank: %i %i" % (xh_loc.shape[1], r_loc))
        r_scale = rank_of_rows(xh_scale, self._groups_scale)
        if r_scale != xh_scale.shape[1]:
            raise ValueError("constrained design matrix design_scale is not full rank: %i %i"
                             % (xh_scale.shape[1], r_scale))
        self.xh_loc, self.xh_scale = xh_loc, xh_scale
        if size_factors is not None:
            sf = np.asarray(size_factors, dtype=np.float64).reshape(-1)
            assert sf.shape[0] == n, "data matrix and size factors must contain same number of cells"
            assert np.all(sf > 0), "size_factors <= 0 found, please remove these cells"
            self.size_factors = np.log(sf).reshape(-1, 1)     # log space, like batchglm (det_cont.py:224)
        else:
            self.size_factors = None
        self.num_observations = n
        self.num_features = shape[1]
        self.features = np.asarray(feature_names) if feature_names is not None \
            else np.asarray(["feature_%d" % i for i in range(shape[1])])
        self.num_loc_params = self.constraints_loc.shape[1]
        self.num_scale_params = self.constraints_scale.shape[1]
        self.num_design_loc_params = self.design_loc.shape[1]
        self.num_design_scale_params = self.design_scale.shape[1]
        self.chunk_size_cells, self.chunk_size_genes, self.cast_dtype = chunk_size_cells, chunk_size_genes, cast_dtype


class Model:
    """batchglm model view: the attributes diffxpy reads (tests.py:779, 802; det.py:561, 619, 1147, 1259-1260)."""

    def __init__(self, estimator):
        self._e = estimator

    @property
    def a_var(self):
        return self._e.a_var

    @property
    def b_var(self):
        return self._e.b_var

    @property
    def a(self):
        return self._e.input_data.constraints_loc @ self._e.a_var

    @property
    def b(self):
        return self._e.input_data.constraints_scale @ self._e.b_var

    @property
    def design_loc(self):
        return self._e.input_data.design_loc

    @property
    def design_scale(self):
        return self._e.input_data.design_scale

    @property
    def design_loc_names(self):
        return self._e.input_data.loc_names

    @property
    def design_scale_names(self):
        return self._e.input_data.scale_names

    @property
    def constraints_loc(self):
        return self._e.input_data.constraints_loc

    @property
    def constraints_scale(self):
        return self._e.input_data.constraints_scale

    @staticmethod
    def inverse_link_loc(x):
        return np.exp(x)

    @staticmethod
    def inverse_link_scale(x):
        return np.exp(x)

    @staticmethod
    def link_loc(x):
        return np.log(x)

    @property
    def eta_loc(self):
        eta = self.design_loc @ self.a
        if self._e.input_data.size_factors is not None:
            eta = eta + self._e.input_data.size_factors
        return np.clip(eta, _ETA_MIN, _ETA_MAX)

    @property
    def location(self):
        return np.exp(self.eta_loc)

    @property
    def scale(self):
        return np.exp(np.clip(self.design_scale @ self.b, _ETA_MIN, _ETA_MAX))

    @property
    def x(self):
        return self._e.x


_ETA_MIN = np.log(np.nextafter(0.0, 1.0)) / 2.5
_ETA_MAX = np.nextafter(np.log(np.finfo(np.float64).max), -np.inf) / 2.5


class Estimator:
    """batchglm numpy ``glm_nb`` Estimator protocol on the B200 (constructed at tests.py:222-229)."""

    def __init__(self, input_data, init_a="AUTO", init_b="AUTO", dtype="float64", quick_scale=False,
                 batch_size=None, device=None, distributed="auto", **kwargs):
        if dtype not in ("float64", np.float64):
            logger.warning("dtype %s requested: the CUDA path always computes in float64", dtype)
        self.input_data = input_data
        self._init_a, self._init_b = init_a, init_b
        self._quick_scale = bool(quick_scale)
        self._device = device
        self._distributed = distributed
        self.model = Model(self)
        self._res = None
        self._dev = None
        self._plan = None
        self._timings = {}

    # ------------------------------------------------------------------------------------------
    def initialize(self):
        """batchglm init_par decisions + upload of the gene shard and the design groups."""
        idata = self.input_data
        dev = cuda_device(self._device)
        self._dev = dev
        n = idata.num_observations
        g_total = idata.num_features
        # gene shard of this rank (contiguous gene ranges, no collective on the data path)
        self._world = dxdist.world(self._distributed)
        # (equal non-zero counts for sparse input, CSR or CSC; every rank computes the same cuts from the host matrix)
        self._gene_ranges = dxdist.gene_ranges(g_total, self._world.size,
                                               dxdist.gene_weights(idata.data) if self._world.size > 1 else None)
        lo, hi = self._gene_ranges[self._world.rank]
        self._gene_range = (lo, hi)
        # (the shard itself is taken over at the END of this method: the host work below -- design grouping, rank checks,
        # pseudo-inverse -- runs while the matrix still uploads in the background, device.prefetch)
        # ---- resolve init modes (batchglm init_par) -------------------------------------------------
        init_a, init_b = self._init_a, self._init_b
        a_given = b_given = None
        sf_nontrivial = idata.size_factors is not None and np.any(idata.size_factors != 0.0)
        train_loc = True
        if isinstance(init_a, str):
            mode = init_a.lower()
            if mode == "auto":
                # (batchglm: the design holds exactly the values 0 and 1 -- read off the distinct rows when they are known,
                # else with three comparisons instead of a sort of all entries)
                d = idata._groups_loc[0] if (idata._groups_loc is not None and idata.xh_loc is idata.design_loc) \
                    else idata.design_loc
                is0, is1 = d == 0.0, d == 1.0
                one_hot = bool(np.all(is0 | is1) and is0.any() and is1.any())
                mode = "closed_form" if one_hot else "standard"
            if mode not in ("closed_form", "standard", "all_zero"):
                raise ValueError("init_a string %s not recognized" % init_a)
            a_mode = {"closed_form": _lib.DX_INIT_CLOSED_FORM, "standard": _lib.DX_INIT_STANDARD,
                      "all_zero": _lib.DX_INIT_ALL_ZERO}[mode]
        else:
            a_given = np.ascontiguousarray(np.asarray(init_a, dtype=np.float64))
            if a_given.shape != (idata.num_loc_params, g_total):
                raise ValueError("init_a has shape %s, expected %s" % (a_given.shape, (idata.num_loc_params, g_total)))
            a_mode = _lib.DX_INIT_GIVEN
        if isinstance(init_b, str):
            mode_b = init_b.lower()
            if mode_b == "auto":
                mode_b = "standard"
            if mode_b not in ("standard", "all_zero", "closed_form"):
                raise ValueError("init_b string %s not recognized" % init_b)
            # closed form (batchglm closedform_nb_glm_logphi; the reference's two-group wrappers pass it, tests.py:1067-1068,
            # 1245-1246): resolved on the first train() call from group-wise moments into given start values
            b_mode = {"standard": _lib.DX_INIT_STANDARD, "all_zero": _lib.DX_INIT_ALL_ZERO,
                      "closed_form": _B_CLOSED_FORM}[mode_b]
        else:
            b_given = np.ascontiguousarray(np.asarray(init_b, dtype=np.float64))
            if b_given.shape != (idata.num_scale_params, g_total):
                raise ValueError("init_b has shape %s, expected %s" % (b_given.shape, (idata.num_scale_params, g_total)))
            b_mode = _lib.DX_INIT_GIVEN
        # ---- design groups -----------------------------------------------------------------------
        cols = [idata.xh_loc, idata.xh_scale]
        if idata.size_factors is not None:
            cols.append(idata.size_factors.reshape(-1, 1))
        grouped = False
        joined = None
        if idata.size_factors is None and idata._groups_loc is not None and idata._groups_scale is not None:
            # both designs were grouped at construction: the joint grouping follows from the two index maps
            joined = join_groupings([idata._groups_loc, idata._groups_scale])
        if joined is not None:
            uniq, inv = joined
            grouped = uniq.shape[0] <= _lib.DX_MAX_GROUPS
        else:
            joint = np.concatenate(cols, axis=1)
            probe = group_rows(joint[:8192])[0].shape[0]
            if probe <= _lib.DX_MAX_GROUPS:
                uniq, inv = group_rows(joint)
                grouped = uniq.shape[0] <= _lib.DX_MAX_GROUPS
        plan = {"a_mode": a_mode, "b_mode": b_mode, "a_given": a_given, "b_given": b_given}
        p, ps = idata.num_loc_params, idata.num_scale_params
        if grouped:
            k = uniq.shape[0]
            xu_loc = np.ascontiguousarray(uniq[:, :p])
            xu_scale = np.ascontiguousarray(uniq[:, p:p + ps])
            offset = np.ascontiguousarray(uniq[:, p + ps]) if idata.size_factors is not None else None
            n_k = np.bincount(inv, minlength=k).astype(np.float64)
            # distinct rows of the raw location design (batchglm closed form works on design_loc, not on groups)
            if idata._groups_loc is not None and idata.xh_loc is idata.design_loc:
                loc_u, loc_inv_cells = idata._groups_loc          # (no constraints: the rows grouped at construction)
            else:
                loc_u, loc_inv_cells = group_rows(idata.design_loc)
            first_cell = np.zeros(k, dtype=np.int64)
            first_cell[inv[::-1]] = np.arange(n - 1, -1, -1)
            loc_group = loc_inv_cells[first_cell].astype(np.int32)
            ud = loc_u @ idata.constraints_loc
            pinv = np.ascontiguousarray(np.linalg.pinv(ud))                    # [p][KL]
            if a_mode == _lib.DX_INIT_CLOSED_FORM:
                # lstsq residuals are empty when the system is square / under-determined (saturated design)
                saturated = loc_u.shape[0] <= p or np.linalg.matrix_rank(ud) < p
                train_loc = (not saturated) or sf_nontrivial
            plan.update(kind="grouped", k=k, kl=loc_u.shape[0], joint_rows=uniq,
                        scale_const=bool(ps == 1 and np.all(xu_scale == xu_scale[0, 0])),
                        cell_group=torch.from_numpy(inv.astype(np.uint8)).to(dev),
                        xu_loc=torch.from_numpy(xu_loc).to(dev), xu_scale=torch.from_numpy(xu_scale).to(dev),
                        offset=None if offset is None else torch.from_numpy(offset).to(dev),
                        n_k=torch.from_numpy(n_k).to(dev), pinv=torch.from_numpy(pinv).to(dev),
                        loc_group=torch.from_numpy(loc_group).to(dev))
        else:
            plan.update(kind="percell",
                        design_loc=torch.from_numpy(np.ascontiguousarray(idata.xh_loc.T)).to(dev),
                        design_scale=torch.from_numpy(np.ascontiguousarray(idata.xh_scale.T)).to(dev),
                        log_sf=None if idata.size_factors is None
                        else torch.from_numpy(np.ascontiguousarray(idata.size_factors.reshape(-1))).to(dev))
            if a_mode == _lib.DX_INIT_CLOSED_FORM:
                loc_u, loc_inv_cells = group_rows(idata.design_loc)
                if loc_u.shape[0] > 500:
                    raise ValueError("large least-square problem in init, likely defined a numeric predictor as categorical")
                plan.update(loc_u=loc_u, loc_inv_cells=loc_inv_cells)
        dxdevice._trace("initialize(): design groups uploaded")
        with torch.cuda.device(dev):
            self.shard = dxdevice.take(idata.data, dev, None if self._world.size == 1 else (lo, hi))
        dxdevice._trace("initialize() done (shard taken over)")
        plan["train_loc"] = bool(train_loc)
        plan["train_scale"] = not self._quick_scale
        self._plan = plan
        self._train_loc, self._train_scale = plan["train_loc"], plan["train_scale"]
        logger.debug("training location model: %s; training scale model: %s; path: %s",
                     self._train_loc, self._train_scale, plan["kind"])

    # ------------------------------------------------------------------------------------------
    def train_sequence(self, training_strategy="DEFAULT", **kwargs):
        if callable(training_strategy):
            return training_strategy(self)
        if isinstance(training_strategy, str):
            key = training_strategy.upper()
            if key == "AUTO":
                key = "DEFAULT"
            if key not in TRAINING_STRATEGIES:
                raise ValueError("training strategy %s not recognized" % training_strategy)
            training_strategy = TRAINING_STRATEGIES[key]
        if training_strategy is None:
            training_strategy = TRAINING_STRATEGIES["DEFAULT"]
        for d in training_strategy:
            d = dict([(k, v) for k, v in d.items() if k not in kwargs])
            self.train(**d, **kwargs)

    def train(self, max_steps=1000, update_b_freq=5, method_b="brent", ftol_b=1e-6, lr_b=1e-2, max_iter_b=1000,
              nproc=3, batch_size=None, **kwargs):
        """One ``train`` call of the numpy estimator = one launch of K_ingest (first call) + K_fit."""
        if self._plan is None:
            self.initialize()
        plan, idata, dev = self._plan, self.input_data, self._dev
        p, ps = idata.num_loc_params, idata.num_scale_params
        shard = self.shard
        g = shard.n_genes
        lo, hi = self._gene_range
        with torch.cuda.device(dev):
            first = self._res is None
            if first:
                a_var = torch.empty((p, g), dtype=torch.float64, device=dev)
                b_var = torch.empty((ps, g), dtype=torch.float64, device=dev)
                if plan["a_given"] is not None:
                    a_var.copy_(torch.from_numpy(np.ascontiguousarray(plan["a_given"][:, lo:hi])))
                if plan["b_given"] is not None:
                    b_var.copy_(torch.from_numpy(np.ascontiguousarray(plan["b_given"][:, lo:hi])))
                a_mode, b_mode = plan["a_mode"], plan["b_mode"]
            else:
                a_var, b_var = self._res["a_var"], self._res["b_var"]
                a_mode = b_mode = _lib.DX_INIT_GIVEN
            fim_inv = torch.empty((g, p, p), dtype=torch.float64, device=dev)
            ll = torch.empty(g, dtype=torch.float64, device=dev)
            jac = torch.empty(g, dtype=torch.float64, device=dev)
            niter = torch.empty(g, dtype=torch.int32, device=dev)
            flags = torch.empty(g, dtype=torch.int32, device=dev)
            opts = _lib.DxFitOpts(max_steps=int(max_steps), update_b_freq=int(update_b_freq),
                                  train_loc=int(plan["train_loc"]), train_scale=int(plan["train_scale"]),
                                  init_a=a_mode, init_b=max(b_mode, 0), newton_max_iter=100, reserved=0,
                                  lltol=LLTOL_BY_FEATURE, newton_xtol=1e-8)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            if plan["kind"] == "grouped":
                if "suff" not in plan:
                    suff = shard.group_suffstats(plan["cell_group"], plan["k"])
                    if b_mode == _B_CLOSED_FORM:
                        b_var.copy_(self._closed_form_b_grouped(suff))
                    plan["suff"] = suff.pack_overflow()
                if b_mode == _B_CLOSED_FORM:
                    opts.init_b = b_mode = _lib.DX_INIT_GIVEN
                if plan.get("scale_const"):
                    opts.reserved |= _lib.DX_OPT_SCALE_CONST          # constant dispersion: register-resident K_fit
                suff = plan["suff"]
                _lib.call("dx_nb_fit_grouped_packed", g, plan["k"], p, ps, ptr(plan["xu_loc"]), ptr(plan["xu_scale"]),
                          ptr(plan["offset"]), ptr(plan["n_k"]), ptr(plan["pinv"]), ptr(plan["loc_group"]),
                          int(plan["kl"]), ptr(suff.tail), ptr(suff.ovf_count), ptr(shard.indptr),
                          ptr(suff.ovf_val), ptr(suff.ovf_group), ptr(suff.ovf_packed), opts, ptr(a_var), ptr(b_var),
                          ptr(fim_inv),
                          ptr(ll), ptr(jac), ptr(niter), ptr(flags), stream_ptr())
            else:
                if first:
                    self._percell_init(a_var, b_var, a_mode, b_mode)
                    opts.init_a = opts.init_b = _lib.DX_INIT_GIVEN
                    b_mode = _lib.DX_INIT_GIVEN
                if ps == 1 and np.all(idata.xh_scale == idata.xh_scale[0, 0]):
                    opts.reserved |= _lib.DX_OPT_SCALE_CONST          # constant dispersion model
                _lib.call("dx_nb_fit_percell", g, shard.n_cells, p, ps, ptr(plan["design_loc"]),
                          ptr(plan["design_scale"]), ptr(plan["log_sf"]), ptr(shard.indptr), ptr(shard.rows),
                          ptr(shard.vals), opts, ptr(a_var), ptr(b_var), ptr(fim_inv), ptr(ll), ptr(jac),
                          ptr(niter), ptr(flags), stream_ptr())
            ev1.record()
            self._timings["train_events"] = (ev0, ev1)
            if first:
                self._niter_total = niter
            else:
                self._niter_total = self._niter_total + niter
            self._res = dict(a_var=a_var, b_var=b_var, fim_inv=fim_inv, ll=ll, jac=jac, niter=self._niter_total,
                             flags=flags)
            self._host = None

    def _percell_init(self, a_var, b_var, a_mode, b_mode):
        """Start values of the per-cell path (batchglm init_par) from device column moments."""
        idata, plan, shard = self.input_data, self._plan, self.shard
        n = idata.num_observations
        dev = self._dev
        inv_sf = None
        if plan["log_sf"] is not None:
            inv_sf = torch.exp(-plan["log_sf"])
        if a_mode == _lib.DX_INIT_STANDARD:
            s1, _ = shard.gene_moments(None)
            a_var.zero_()
            a_var[0] = torch.clamp(torch.log(s1 / n), _ETA_MIN, _ETA_MAX)
        elif a_mode == _lib.DX_INIT_ALL_ZERO:
            a_var.zero_()
        elif a_mode == _lib.DX_INIT_CLOSED_FORM:
            loc_u, inv_cells = plan["loc_u"], plan["loc_inv_cells"]
            kl = loc_u.shape[0]
            means = torch.empty((kl, shard.n_genes), dtype=torch.float64, device=dev)
            for l in range(kl):
                w = torch.from_numpy((inv_cells == l).astype(np.float64)).to(dev)
                cnt = float(w.sum().item())
                if inv_sf is not None:
                    w = w * inv_sf
                s1, _ = shard.gene_moments(w.contiguous())
                means[l] = s1 / cnt
            linked = torch.log(means + np.nextafter(0.0, 1.0))
            pinv = torch.from_numpy(np.linalg.pinv(loc_u @ idata.constraints_loc)).to(dev)
            a_var.copy_(torch.clamp(pinv @ linked, _ETA_MIN, _ETA_MAX))
        if b_mode == _lib.DX_INIT_STANDARD:
            s1, s2 = shard.gene_moments(inv_sf)
            m = s1 / n
            v = s2 / n - m * m
            denom = torch.clamp(v - m, min=float(np.sqrt(np.nextafter(0.0, 1.0))))
            b_var.zero_()
            b_var[0] = torch.clamp(torch.log(m * m / denom + np.nextafter(0.0, 1.0)), _ETA_MIN, _ETA_MAX)
        elif b_mode == _lib.DX_INIT_ALL_ZERO:
            b_var.zero_()
        elif b_mode == _B_CLOSED_FORM:
            scale_u, inv_cells = group_rows(idata.design_scale)
            if scale_u.shape[0] > 500:
                raise ValueError("large least-square problem in init, likely defined a numeric predictor as categorical")
            ks = scale_u.shape[0]
            m = torch.empty((ks, shard.n_genes), dtype=torch.float64, device=dev)
            e2 = torch.empty_like(m)
            for l in range(ks):
                sel = torch.from_numpy((inv_cells == l).astype(np.float64)).to(dev)
                cnt = float(sel.sum().item())
                w = sel if inv_sf is None else sel * inv_sf
                s1, s2 = shard.gene_moments(w.contiguous())          # sums of (x w) and (x w)^2: w is 0 outside the group
                m[l], e2[l] = s1 / cnt, s2 / cnt
            b_var.copy_(self._closed_form_b_solve(m, e2, scale_u))

    def _closed_form_b_solve(self, m, e2, scale_u):
        """batchglm closedform_nb_glm_logphi (oracle/nb_glm.py:init_par): log method-of-moments size of every distinct
        dispersion-design row, pushed through the least-squares solution of those rows."""
        v = e2 - m * m
        denom = torch.clamp(v - m, min=float(np.sqrt(np.nextafter(0.0, 1.0))))
        linked = torch.log(m * m / denom)
        pinv = torch.from_numpy(np.linalg.pinv(scale_u @ self.input_data.constraints_scale)).to(m.device)
        # lstsq of the reference = pseudo-inverse; -inf entries (all-zero groups) propagate like there and are clipped below
        b0 = pinv @ torch.nan_to_num(linked, neginf=_ETA_MIN * 4)
        return torch.clamp(b0, _ETA_MIN, _ETA_MAX)

    def _closed_form_b_grouped(self, suff):
        """Start values of the dispersion model from the sufficient statistics of K_ingest (before K_pack): per design
        group sum(x) and sum(x^2) from the tail counts (+ overflow entries), pooled over the groups that share a
        dispersion-design row, divided by the group's size factor."""
        idata, plan = self.input_data, self._plan
        dev = suff.tail.device
        k, g = plan["k"], suff.tail.shape[0]
        p, ps = idata.num_loc_params, idata.num_scale_params
        tail = suff.tail.to(torch.float64)                                            # [G][K][64]
        j = torch.arange(_lib.DX_HIST_BINS, dtype=torch.float64, device=dev)
        s1 = tail.sum(dim=2)                                                          # sum_j c(j) = sum of counts
        s2 = (tail * (2.0 * j + 1.0)).sum(dim=2)                                      # sum_j c(j) (2 j + 1) = sum of squares
        n_ovf = suff.ovf_count.to(torch.int64)
        if int(n_ovf.sum().item()) > 0:
            genes = torch.nonzero(n_ovf).reshape(-1)
            cnts = n_ovf[genes]
            gene_of = torch.repeat_interleave(genes, cnts)
            start = torch.repeat_interleave(suff.shard.indptr[genes], cnts)
            first = torch.cumsum(cnts, 0) - cnts
            pos = start + (torch.arange(gene_of.shape[0], device=dev) - torch.repeat_interleave(first, cnts))
            val = suff.ovf_val[pos].to(torch.float64)
            flat = gene_of * k + suff.ovf_group[pos].to(torch.int64)
            s1.view(-1).index_add_(0, flat, val)
            s2.view(-1).index_add_(0, flat, val * val)
        # dispersion-design groups: distinct rows of design_scale, numbered through the joint design groups
        joint = plan["joint_rows"]
        scale_rows = joint[:, p:p + ps]
        # design_scale itself (not the constrained one) defines the groups in batchglm; with constraints the distinct
        # rows of the constrained design are a coarsening that spans the same space
        scale_u, grp = group_rows(np.ascontiguousarray(scale_rows))
        n_k = plan["n_k"].to(torch.float64)
        if plan["offset"] is not None:
            isf = torch.exp(-plan["offset"])
            s1, s2 = s1 * isf[None, :], s2 * (isf * isf)[None, :]
        sel = torch.zeros((k, scale_u.shape[0]), dtype=torch.float64, device=dev)
        sel[torch.arange(k, device=dev), torch.from_numpy(grp).to(dev)] = 1.0
        cnt = n_k @ sel                                                               # cells per dispersion group
        m = (s1 @ sel / cnt[None, :]).t().contiguous()                                # [KS][G]
        e2 = (s2 @ sel / cnt[None, :]).t().contiguous()
        v = e2 - m * m
        denom = torch.clamp(v - m, min=float(np.sqrt(np.nextafter(0.0, 1.0))))
        linked = torch.log(m * m / denom)
        pinv = torch.from_numpy(np.linalg.pinv(scale_u)).to(dev)
        b0 = pinv @ torch.nan_to_num(linked, neginf=_ETA_MIN * 4)
        return torch.clamp(b0, _ETA_MIN, _ETA_MAX)

    # ------------------------------------------------------------------------------------------
    def finalize(self):
        """Results to the host (and, on several GPUs, gathered over NCCL so every rank holds all genes)."""
        if self._res is None:
            raise RuntimeError("finalize() called before train_sequence()")
        r = self._res
        with torch.cuda.device(self._dev):
            packed = dxdist.gather_genes(self._world, {
                "a_var": r["a_var"].t().contiguous(), "b_var": r["b_var"].t().contiguous(),
                "fim_inv": r["fim_inv"], "ll": r["ll"], "jac": r["jac"],
                "niter": r["niter"], "flags": r["flags"]}, self.input_data.num_features,
                counts=[b - a for a, b in self._gene_ranges])
        self._host = dict(
            a_var=np.ascontiguousarray(packed["a_var"].cpu().numpy().T),
            b_var=np.ascontiguousarray(packed["b_var"].cpu().numpy().T),
            fisher_inv=packed["fim_inv"].cpu().numpy(),
            log_likelihood=packed["ll"].cpu().numpy(),
            jacobian=packed["jac"].cpu().numpy(),
            niter=packed["niter"].cpu().numpy(),
            error_codes=packed["flags"].cpu().numpy(),
        )
        self._gathered = packed
        return self

    @property
    def device_train_ms(self):
        """Device time of the last train() call (K_ingest on first use + K_fit), from CUDA events."""
        ev = self._timings.get("train_events")
        if ev is None:
            return None
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1])

    def _h(self, key):
        if getattr(self, "_host", None) is None:
            self.finalize()
        return self._host[key]

    # ---- attributes diffxpy reads ----------------------------------------------------------------
    @property
    def a_var(self):
        return self._h("a_var")

    @property
    def b_var(self):
        return self._h("b_var")

    @property
    def fisher_inv(self):
        return self._h("fisher_inv")

    @property
    def hessian(self):
        with np.errstate(all="ignore"):
            return -np.linalg.inv(self.fisher_inv)

    @property
    def log_likelihood(self):
        return self._h("log_likelihood")

    @property
    def jacobian(self):
        return self._h("jacobian")

    @property
    def niter(self):
        return self._h("niter")

    @property
    def error_codes(self):
        return self._h("error_codes")

    @property
    def loss(self):
        return np.sum(self.log_likelihood)

    @property
    def x(self):
        return self.input_data.x

    @property
    def X(self):
        return self.input_data.x

    @property
    def features(self):
        return self.input_data.features

    @property
    def location(self):
        return self.model.location

    @property
    def scale(self):
        return self.model.scale

    @property
    def par_link_loc(self):
        return self.model.a

    @property
    def par_link_scale(self):
        return self.model.b

    def gene_means(self):
        """Column means of the raw counts (det.py:775-781) from one streaming pass on the device."""
        with torch.cuda.device(self._dev):
            s1, _ = self.shard.gene_moments(None)
            out = dxdist.gather_genes(self._world, {"m": s1 / self.input_data.num_observations},
                                      self.input_data.num_features, counts=[b - a for a, b in self._gene_ranges])
        return out["m"].cpu().numpy()
