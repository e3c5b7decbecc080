"""
Formula -> design matrix and constraint helpers without patsy.

Mirrors the helpers diffxpy relays to ``batchglm.data`` (/root/reference/diffxpy/testing/utils.py:127-319:
``design_matrix``, ``preview_coef_names``, ``constraint_system_from_star``, ``constraint_matrix_from_string``,
``constraint_matrix_from_dict``, ``view_coef_names``).  patsy is not a dependency here, so the formula
language is the additive subset the reference's tests use:

    "~ 1 + condition + batch"   "~ condition"   "~ 0 + group"   "~ 1 + condition + numeric1"   "~ 1"   "a:b"

Categorical terms are treatment coded with sorted levels (first level absorbed when an intercept or an earlier
categorical term spans it), columns are named like patsy: ``Intercept``, ``condition[T.1]``, ``group[a]``.
"""
import itertools

import numpy as np
import pandas as pd


class DesignInfo:
    """The slice of patsy.DesignInfo that diffxpy touches (tests.py:675-678, det.py:540-553)."""

    def __init__(self, column_names, term_names, term_slices):
        self.column_names = list(column_names)
        self.term_names = list(term_names)
        self.term_name_slices = dict(zip(self.term_names, term_slices))
        self.term_slices = self.term_name_slices

    def slice(self, name):
        if name in self.term_name_slices:
            return self.term_name_slices[name]
        if name in self.column_names:
            i = self.column_names.index(name)
            return slice(i, i + 1)
        raise ValueError("unknown term or column name %r" % (name,))

    def subset(self, terms):
        if isinstance(terms, str):
            terms = [terms]
        terms = [t for t in self.term_names if t in set(terms)]
        cols, slices, pos = [], [], 0
        for t in terms:
            sl = self.term_name_slices[t]
            names = self.column_names[sl]
            cols.extend(names)
            slices.append(slice(pos, pos + len(names)))
            pos += len(names)
        return DesignInfo(cols, terms, slices)


class DesignMatrix(np.ndarray):
    """numpy array carrying ``design_info`` (stands in for patsy.design_info.DesignMatrix)."""

    def __new__(cls, array, design_info):
        obj = np.asarray(array, dtype=np.float64).view(cls)
        obj.design_info = design_info
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.design_info = getattr(obj, "design_info", None)


def _parse_formula(formula):
    rhs = formula.split("~", 1)[1] if "~" in formula else formula
    rhs = rhs.replace("-", "+-")
    intercept = True
    terms = []
    for tok in rhs.split("+"):
        tok = tok.strip().replace(" ", "")
        if tok == "":
            continue
        if tok in ("1",):
            intercept = True
        elif tok in ("0", "-1"):
            intercept = False
        elif tok.startswith("-"):
            raise ValueError("cannot remove term %r: only '- 1' is supported" % tok)
        elif "*" in tok:
            parts = [t.strip() for t in tok.split("*")]
            if len(parts) != 2:
                raise ValueError("only two-way '*' terms are supported: %r" % tok)
            for t in (parts[0], parts[1], parts[0] + ":" + parts[1]):
                if t not in terms:
                    terms.append(t)
        else:
            if tok not in terms:
                terms.append(tok)
    return intercept, terms


def _is_categorical(sample_description, col, as_categorical):
    if isinstance(as_categorical, bool):
        flag = as_categorical
    else:
        flag = bool(list(as_categorical)[list(sample_description.columns).index(col)])
    if flag:
        return True
    return not pd.api.types.is_numeric_dtype(sample_description[col])


def _levels(series):
    if isinstance(series.dtype, pd.CategoricalDtype):
        # patsy keeps the order of an existing categorical (the first category is the reference level); levels that do
        # not occur are left out here (patsy would add all-zero columns, which the rank check of de.test.* rejects)
        present = set(pd.unique(series.astype(object)))
        return [lv for lv in series.cat.categories if lv in present]
    vals = pd.unique(series)
    try:
        return sorted(vals)
    except TypeError:
        return sorted(vals, key=str)


def _resolve_factor(sample_description, term):
    """'x' -> ('x', False); patsy's categorical marker 'C(x)' -> ('x', True).  Raises for unknown columns."""
    col, forced = term, False
    if term.startswith("C(") and term.endswith(")"):
        col, forced = term[2:-1], True
    if col not in sample_description.columns:
        raise ValueError("term %r not found in sample description columns %s" % (term, list(sample_description.columns)))
    return col, forced


def _factor_columns(sample_description, term, as_categorical, full_rank):
    """Columns of one main-effect term -> (matrix n x k, names)."""
    col, forced = _resolve_factor(sample_description, term)
    s = sample_description[col]
    if forced or _is_categorical(sample_description, col, as_categorical):
        col = term                      # patsy names the columns after the term as written: C(x)[T.level]
        codes = None
        if isinstance(s.dtype, pd.CategoricalDtype):
            levels = _levels(s)
        else:
            # one hash pass over the column gives codes and distinct values; the levels are the sorted values (_levels), the
            # codes follow by a permutation of the small code range
            raw, uniq = pd.factorize(s, use_na_sentinel=True)
            uniq = list(uniq)
            try:
                order = sorted(range(len(uniq)), key=lambda i: uniq[i])
            except TypeError:
                order = sorted(range(len(uniq)), key=lambda i: str(uniq[i]))
            levels = [uniq[i] for i in order]
            perm = np.empty(len(uniq) + 1, dtype=np.int64)
            perm[np.asarray(order, dtype=np.int64)] = np.arange(len(uniq))
            perm[len(uniq)] = -1                      # the NA sentinel -1 indexes the last entry
            codes = perm[np.asarray(raw, dtype=np.int64)]
        if full_rank:
            use = levels
            names = ["%s[%s]" % (col, lv) for lv in use]
        else:
            use = levels[1:]
            names = ["%s[T.%s]" % (col, lv) for lv in use]
        # one pass over the column: level codes, then a one-hot scatter (comparing the column with every level in turn
        # costs a string comparison pass per level)
        if codes is None:
            codes = pd.Categorical(s, categories=levels).codes.astype(np.int64)
        onehot = np.zeros((len(s), len(levels)), dtype=np.float64)
        ok = codes >= 0
        onehot[np.nonzero(ok)[0], codes[ok]] = 1.0
        mat = onehot if full_rank else onehot[:, 1:]
        return mat, names
    return np.asarray(s.values, dtype=np.float64)[:, None], [col]


def design_matrix(sample_description=None, formula=None, as_categorical=True, dmat=None, return_type="patsy"):
    """batchglm.data.design_matrix (relayed at utils.py:127-176)."""
    if sample_description is None and dmat is None:
        raise ValueError("supply either sample_description or dmat")
    if dmat is None and formula is None:
        raise ValueError("supply either dmat or formula")
    if dmat is not None:
        if isinstance(dmat, pd.DataFrame):
            names = [str(c) for c in dmat.columns]
            arr = dmat.values
        elif isinstance(dmat, DesignMatrix):
            names = dmat.design_info.column_names
            arr = np.asarray(dmat)
        else:
            arr = np.asarray(dmat, dtype=np.float64)
            names = ["x%d" % i for i in range(arr.shape[1])]
        di = DesignInfo(names, names, [slice(i, i + 1) for i in range(len(names))])
        out = DesignMatrix(arr, di)
    else:
        sample_description = sample_description.copy()
        intercept, terms = _parse_formula(formula)
        blocks, names, tnames, tslices = [], [], [], []
        n = sample_description.shape[0]
        pos = 0
        if intercept:
            blocks.append(np.ones((n, 1)))
            names.append("Intercept")
            tnames.append("Intercept")
            tslices.append(slice(0, 1))
            pos = 1
        # patsy's coding rule (patsy/build.py, "redundancy"): a term with categorical factors c1..ck stands for the sum of
        # the sub-terms over all subsets of them, each factor reduced-rank (treatment) coded; sub-terms that an earlier term
        # (or the intercept) already covers are dropped, and a remaining pair "S" and "S + f" folds into one sub-term in
        # which f is full-rank coded.  Numeric factors define separate buckets: 'a:x' is only redundant with terms that
        # carry the same numeric factors.
        used = set()
        if intercept:
            used.add((frozenset(), frozenset()))

        def _numeric_factors(term):
            out_ = []
            for f in [x.strip() for x in term.split(":")]:
                col_f, forced_f = _resolve_factor(sample_description, f)
                if not (forced_f or _is_categorical(sample_description, col_f, as_categorical)):
                    out_.append(f)
            return frozenset(out_)

        # patsy's column order: terms without numeric factors first, then one group per set of numeric factors in order
        # of first appearance; inside a group by interaction order, ties in formula order
        buckets = [frozenset()]
        for t in terms:
            b = _numeric_factors(t)
            if b not in buckets:
                buckets.append(b)
        terms = sorted(terms, key=lambda t: (buckets.index(_numeric_factors(t)), t.count(":"), terms.index(t)))
        for t in terms:
            factors = [x.strip() for x in t.split(":")]
            cats, nums = [], []
            for f in factors:
                col_f, forced_f = _resolve_factor(sample_description, f)
                (cats if (forced_f or _is_categorical(sample_description, col_f, as_categorical)) else nums).append(f)
            bucket = frozenset(nums)
            subsets = [frozenset(c for c, keep in zip(cats, bits) if keep)
                       for bits in itertools.product([False, True], repeat=len(cats))]
            subsets.sort(key=lambda st: (len(st), [cats.index(c) for c in cats if c in st]))
            fresh = [st for st in subsets if (st, bucket) not in used]
            used.update((st, bucket) for st in fresh)
            subterms = [{c: False for c in cats if c in st} for st in fresh]        # factor -> full-rank coded?
            merged = True
            while merged:
                merged = False
                for a_i, short in enumerate(subterms):
                    for b_i in range(a_i + 1, len(subterms)):
                        long_ = subterms[b_i]
                        extra = [c for c in long_ if c not in short]
                        if len(long_) == len(short) + 1 and len(extra) == 1 and not long_[extra[0]] \
                                and all(long_[c] == short[c] for c in short):
                            folded = dict(short)
                            folded[extra[0]] = True
                            subterms[b_i] = folded          # (patsy keeps the folded sub-term where the longer one stood)
                            subterms.pop(a_i)
                            merged = True
                            break
                    if merged:
                        break
            num_mat = np.ones((n, 1))
            for f in nums:
                num_mat = num_mat * np.asarray(sample_description[_resolve_factor(sample_description, f)[0]].values,
                                               dtype=np.float64)[:, None]
            t_blocks, t_names = [], []
            for st in subterms:
                mats, nms = [np.ones((n, 1))], [""]
                for f in factors:                     # factor order of the term; the first factor varies fastest
                    if f in nums:
                        continue
                    if f not in st:
                        continue
                    m_f, n_f = _factor_columns(sample_description, f, True, full_rank=st[f])
                    if nms == [""]:                   # first categorical factor of the sub-term: nothing to multiply yet
                        mats, nms = [m_f], list(n_f)
                        continue
                    mats = [np.concatenate([mats[0][:, [a_]] * m_f[:, [b_]] for b_ in range(m_f.shape[1])
                                            for a_ in range(mats[0].shape[1])], axis=1)]
                    nms = [(x + ":" if x else "") + y for y in n_f for x in nms]
                mat_st, names_st = (mats[0] * num_mat if nums else mats[0]), nms
                if nums:
                    names_st = [(x + ":" if x else "") + ":".join(nums) for x in names_st]
                if mat_st.shape[1]:
                    t_blocks.append(mat_st)
                    t_names.extend(names_st)
            mat = np.concatenate(t_blocks, axis=1) if t_blocks else np.zeros((n, 0))
            cn = t_names
            blocks.append(mat)
            names.extend(cn)
            tnames.append(t)
            tslices.append(slice(pos, pos + mat.shape[1]))
            pos += mat.shape[1]
        arr = np.concatenate(blocks, axis=1) if blocks else np.zeros((n, 0))
        out = DesignMatrix(arr, DesignInfo(names, tnames, tslices))
    if return_type == "dataframe":
        return pd.DataFrame(np.asarray(out), columns=out.design_info.column_names)
    return out


def preview_coef_names(sample_description, formula, as_categorical=True):
    """batchglm.data.preview_coef_names (relayed at utils.py:179-212)."""
    return design_matrix(sample_description=sample_description.iloc[:].copy(), formula=formula,
                         as_categorical=as_categorical).design_info.column_names


def view_coef_names(dmat):
    """batchglm.data.view_coef_names."""
    if isinstance(dmat, pd.DataFrame):
        return np.asarray(dmat.columns)
    if isinstance(dmat, DesignMatrix):
        return np.asarray(dmat.design_info.column_names)
    raise ValueError("dmat type %s not recognized" % type(dmat))


def _parse_linear_constraint(expr, coef_names):
    """'a + b + c = 0' -> 0/1 indicator over coef_names (order of appearance kept)."""
    if "=" not in expr:
        raise ValueError("constraint %r has no '='" % expr)
    lhs, rhs = expr.split("=", 1)
    if rhs.strip() not in ("0", "0.0"):
        raise ValueError("only '... = 0' constraints are supported: %r" % expr)
    order = []
    for tok in lhs.split("+"):
        tok = tok.strip()
        if tok not in coef_names:
            raise ValueError("constraint coefficient %r not in %s" % (tok, list(coef_names)))
        order.append(list(coef_names).index(tok))
    return order


def constraint_matrix_from_string(dmat, coef_names, constraints):
    """batchglm.data.constraint_matrix_from_string: rows = all parameters, columns = free parameters; the
    first coefficient named in each constraint is the dependent one (-1 on the others of its set)
    (semantics documented at tests.py:54-71)."""
    assert len(constraints) > 0, "supply constraints"
    coef_names = list(coef_names)
    n_par_all = np.asarray(dmat).shape[1]
    n_par_free = n_par_all - len(constraints)
    parsed = [_parse_linear_constraint(c, coef_names) for c in constraints]
    idx_constr = np.asarray([o[0] for o in parsed])
    idx_depending = [o[1:] for o in parsed]
    idx_unconstr = np.asarray(sorted(set(range(n_par_all)) - set(idx_constr.tolist())))
    cmat = np.zeros([n_par_all, n_par_free])
    for i in range(n_par_all):
        if i in idx_constr:
            deps = idx_depending[int(np.where(idx_constr == i)[0][0])]
            cols = [int(np.where(idx_unconstr == d)[0][0]) for d in deps]
            cmat[i, cols] = -1
        else:
            cmat[i, int(np.where(idx_unconstr == i)[0][0])] = 1
    return cmat


def string_constraints_from_dict(sample_description, constraints):
    out = []
    for x, y in constraints.items():
        d_x = design_matrix(sample_description, "~ 0 + " + x)
        if y == "1":
            out.append("+".join(d_x.design_info.column_names) + "=0")
            continue
        d_y = design_matrix(sample_description, "~ 0 + " + y)
        for j in range(d_y.shape[1]):
            grouping = np.asarray(d_y)[:, j]
            inside = np.where(np.sum(np.asarray(d_x)[grouping == 1, :], axis=0) > 0)[0]
            outside = np.where(np.sum(np.asarray(d_x)[grouping == 0, :], axis=0) > 0)[0]
            assert len(set(inside.tolist()) & set(outside.tolist())) == 0, \
                "proposed grouping of constraints is not nested, read docstrings"
            out.append("+".join(np.asarray(d_x.design_info.column_names)[inside]) + "=0")
    return out


def constraint_matrix_from_dict(sample_description, formula, as_categorical=True, constraints=None, return_type="patsy"):
    """batchglm.data.constraint_matrix_from_dict: constrained factors are re-coded with all their levels and
    one sum-to-zero constraint is added per group of the grouping factor."""
    assert len(constraints) > 0, "supply constraints"
    intercept, terms = _parse_formula(formula)
    free_terms = [t for t in terms if t not in constraints]
    f_unconstrained = "~ " + ("1" if intercept else "0") + "".join(" + " + t for t in free_terms)
    dmat = design_matrix(sample_description, f_unconstrained, as_categorical)
    arr = np.asarray(dmat)
    coef_names = list(dmat.design_info.column_names)
    term_names = list(dmat.design_info.term_names)
    slices = [dmat.design_info.term_name_slices[t] for t in term_names]
    for x in constraints.keys():
        d_x = design_matrix(sample_description, "~ 0 + " + x)
        slices.append(slice(arr.shape[1], arr.shape[1] + d_x.shape[1]))
        arr = np.hstack([arr, np.asarray(d_x)])
        coef_names.extend(d_x.design_info.column_names)
        term_names.append(x)
    constraints_ls = string_constraints_from_dict(sample_description, constraints)
    out = DesignMatrix(arr, DesignInfo(coef_names, term_names, slices))
    cmat = constraint_matrix_from_string(out, coef_names, constraints_ls)
    if return_type == "dataframe":
        out = pd.DataFrame(np.asarray(out), columns=coef_names)
    return out, coef_names, cmat, term_names


def constraint_system_from_star(dmat=None, sample_description=None, formula=None, as_categorical=True,
                                constraints=None, return_type="patsy"):
    """batchglm.data.constraint_system_from_star (relayed at utils.py:215-273; consumed at tests.py:647-662):
    returns (design, coef_names, constraint_matrix or None, term_names)."""
    if sample_description is None and dmat is None:
        raise ValueError("supply either sample_description or dmat")
    if dmat is None and formula is None:
        raise ValueError("supply either dmat or formula")
    if dmat is not None and isinstance(constraints, dict):
        raise ValueError("dmat was supplied even though constraints were given as dict")
    if isinstance(constraints, dict) and len(constraints) > 0:
        return constraint_matrix_from_dict(sample_description, formula, as_categorical, constraints, return_type)
    dm = design_matrix(sample_description=sample_description, formula=formula, as_categorical=as_categorical,
                       dmat=dmat, return_type="patsy")
    coef_names = list(dm.design_info.column_names)
    term_names = list(dm.design_info.term_names)
    if constraints is None or (isinstance(constraints, dict) and len(constraints) == 0):
        cmat = None
    elif isinstance(constraints, (list, tuple)):
        cmat = constraint_matrix_from_string(dm, coef_names, list(constraints))
    elif isinstance(constraints, np.ndarray):
        cmat = constraints
    else:
        raise ValueError("constraint format %s not recognized" % type(constraints))
    if return_type == "dataframe":
        dm = pd.DataFrame(np.asarray(dm), columns=coef_names)
    return dm, coef_names, cmat, term_names


def bin_continuous_covariate(sample_description, factor_to_bin, bins):
    """batchglm.data.bin_continuous_covariate (relayed at utils.py:276-319): quantile bins."""
    bins = np.asarray(bins)
    v = np.asarray(sample_description[factor_to_bin].values, dtype=np.float64)
    edges = np.quantile(v, bins)
    return np.asarray([edges[np.searchsorted(edges, x, side="right") - 1] for x in v])
