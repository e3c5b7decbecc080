"""
diffxpy_b200 -- B200 (sm_100a) native implementation of diffxpy's per-gene negative-binomial GLM fit and
Wald / LRT tests and of its model-free Welch-t / Wilcoxon tests, behind diffxpy's own ``de.test.*`` API.

    import diffxpy_b200.api as de
    res = de.test.wald(data=csr_counts, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                       sample_description=obs, gene_names=names)
    res.summary()
"""
import logging

__version__ = "0.1.0"
logging.getLogger("diffxpy").addHandler(logging.NullHandler())
