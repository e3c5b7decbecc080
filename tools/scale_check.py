#!/usr/bin/env python
"""
Full-size checks of the BASELINE.json configs that bench.py does not time (configs[2], [3], [4]): wall time of
the device path plus size-independent properties (rank-sum identity, scipy spot checks, convergence, LRT sign).
Writes one JSON object per config to stdout.     python tools/scale_check.py [cfg3] [cfg4] [cfg5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffxpy_b200 import _lib  # noqa: E402
from diffxpy_b200.testing.tests import _fit  # noqa: E402
from diffxpy_b200.stats.stats import two_group_device  # noqa: E402
import pandas as pd  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return out, float(np.median(ts))


def cfg5(n_cells=1_000_000, n_genes=2500):
    """rank_test / t_test, 1M cells, two groups (one GPU's 1/8 share of 20k genes)."""
    import scipy.stats
    shard, cond, _ = bench.make_shard(torch, dev, seed=5, n_genes=n_genes, n_cells=n_cells, chunk=100)
    n0, n1 = int((cond == 0).sum()), int((cond == 1).sum())
    res, sec = timed(lambda: two_group_device(shard, cond.astype(np.uint8), n0, n1, do_rank=True))
    u1 = res["u1x2"] / 2.0
    # spot check three genes against scipy on the host
    indptr = shard.indptr.cpu().numpy()
    worst_p = 0.0
    for g in (0, n_genes // 2, n_genes - 1):
        sl = slice(indptr[g], indptr[g + 1])
        col = np.zeros(n_cells, dtype=np.float64)
        col[shard.rows[sl].cpu().numpy()] = shard.vals[sl].cpu().numpy()
        ref = scipy.stats.mannwhitneyu(col[cond == 0], col[cond == 1], use_continuity=True, alternative="two-sided")
        assert ref.statistic == u1[g], "U1 differs from scipy at gene %d" % g
        worst_p = max(worst_p, abs(ref.pvalue - res["out"][g, _lib.TG_PVAL_RANK]) / max(ref.pvalue, 1e-300))
    return {"config": "configs[4]: rank_test + t_test, %d cells x %d genes (1/8 of 20k), two groups" % (n_cells, n_genes),
            "nnz": shard.nnz, "seconds": sec, "genes_per_s": n_genes / sec,
            "stream_gbs": 8 * shard.nnz / sec / 1e9,
            "u1_matches_scipy_exactly": True, "max_rel_dp_vs_scipy": worst_p,
            "u_in_range": bool(np.all((u1 >= 0) & (u1 <= float(n0) * n1))),
            "flags_nonzero": int(np.count_nonzero(res["flags"]))}


def ranklog(n_cells=100_000, n_genes=2_000):
    """rank test on log-normalised data (every non-zero is a non-count value): K_ingest + K_twogroup + the workspace sort."""
    import scipy.stats
    from diffxpy_b200.device import CountShard
    bench.MU_HI = 3.0
    shard, cond, _ = bench.make_shard(torch, dev, seed=6, n_genes=n_genes, n_cells=n_cells, chunk=200)
    bench.MU_HI = 1.0
    sf = torch.rand(n_cells, device=dev) + 0.5
    vals = torch.log1p(shard.vals / sf[shard.rows.long()])
    shard = CountShard(shard.indptr, shard.rows, vals.contiguous(), n_cells)
    n0, n1 = int((cond == 0).sum()), int((cond == 1).sum())
    res, sec = timed(lambda: two_group_device(shard, cond.astype(np.uint8), n0, n1, do_rank=True))
    u1 = res["u1x2"] / 2.0
    indptr = shard.indptr.cpu().numpy()
    nnz_g = np.diff(indptr)
    worst_p = 0.0
    for g in (int(np.argmax(nnz_g)), n_genes // 2, int(np.argmin(nnz_g))):
        sl = slice(indptr[g], indptr[g + 1])
        col = np.zeros(n_cells, dtype=np.float64)
        col[shard.rows[sl].cpu().numpy()] = shard.vals[sl].cpu().numpy()
        ref = scipy.stats.mannwhitneyu(col[cond == 0], col[cond == 1], use_continuity=True, alternative="two-sided")
        assert ref.statistic == u1[g], "U1 differs from scipy at gene %d" % g
        worst_p = max(worst_p, abs(ref.pvalue - res["out"][g, _lib.TG_PVAL_RANK]) / max(ref.pvalue, 1e-300))
    return {"config": "rank_test on log1p(count / size factor), %d cells x %d genes, two groups" % (n_cells, n_genes),
            "nnz": shard.nnz, "max_nnz_per_gene": int(nnz_g.max()), "median_nnz_per_gene": float(np.median(nnz_g)),
            "seconds": sec, "genes_per_s": n_genes / sec, "u1_matches_scipy_exactly": True,
            "max_rel_dp_vs_scipy": worst_p, "flags_nonzero": int(np.count_nonzero(res["flags"]))}


def cfg3(n_cells=100_000, n_genes=20_000):
    """LRT: full '~1+condition+batch' vs reduced '~1+batch', scale model '~1+batch' (8 dispersion coefficients)."""
    shard, cond, batch = bench.make_shard(torch, dev, seed=3, n_genes=n_genes, n_cells=n_cells)
    xd_full = bench.design_for(cond, batch)
    xd_red = np.delete(xd_full, 1, axis=1)
    zd = np.delete(xd_full, 1, axis=1)
    names = ["g%d" % i for i in range(n_genes)]
    df = lambda m, pre: pd.DataFrame(m, columns=["%s%d" % (pre, i) for i in range(m.shape[1])])

    def run():
        red = _fit("nb", shard, df(xd_red, "r"), df(zd, "s"), gene_names=names)
        full = _fit("nb", shard, df(xd_full, "f"), df(zd, "s"), gene_names=names)
        return red, full
    (red, full), sec = timed(run, reps=2)
    from diffxpy_b200.stats import stats as dstats
    pv = dstats.likelihood_ratio_test(full.log_likelihood, red.log_likelihood, 17, 16)
    conv = float(np.mean((full.error_codes & 1) != 0))
    return {"config": "configs[2]: de.test.lrt, %d x %d, full p_loc 9 vs reduced 8, scale '~1+batch' (8 coefs)" % (n_cells, n_genes),
            "nnz": shard.nnz, "seconds_two_fits": sec, "genes_per_s": n_genes / sec,
            "frac_converged_full": conv, "mean_niter_full": float(np.mean(full.niter)),
            "ll_full_ge_ll_reduced_frac": float(np.mean(full.log_likelihood >= red.log_likelihood - 1e-6 * np.abs(red.log_likelihood))),
            "pval_finite_frac": float(np.mean(np.isfinite(pv)))}


def cfg4(n_cells=int(os.environ.get("DXB200_CFG4_CELLS", "250000")), n_genes=int(os.environ.get("DXB200_CFG4_GENES", "9472"))):
    """Wald with a continuous covariate: cubic B-spline basis (11 + intercept), per-cell size factors -> per-cell path.
    9472 genes = one full wave of 32-gene tiles, two CTAs on each of 148 SMs (configs[3] has 30k genes = 3.2 waves)."""
    from scipy.interpolate import BSpline
    shard, cond, batch = bench.make_shard(torch, dev, seed=4, n_genes=n_genes, n_cells=n_cells, chunk=200)
    rng = np.random.default_rng(4)
    t = rng.uniform(0, 1, n_cells)
    df_, deg = 11, 3
    inner = np.quantile(t, np.linspace(0, 1, df_ - deg + 1)[1:-1])
    knots = np.concatenate([[0.0] * (deg + 1), inner, [1.0] * (deg + 1)])
    basis = BSpline.design_matrix(np.clip(t, 0, 1 - 1e-12), knots, deg).toarray()[:, 1:]       # drop one column: intercept present
    xd = np.concatenate([np.ones((n_cells, 1)), basis], axis=1)
    sf = rng.uniform(0.5, 1.5, n_cells)
    names = ["g%d" % i for i in range(n_genes)]
    cols = ["Intercept"] + ["bs%d" % i for i in range(basis.shape[1])]

    def run():
        return _fit("nb", shard, pd.DataFrame(xd, columns=cols), pd.DataFrame(np.ones((n_cells, 1)), columns=["Intercept"]),
                    gene_names=names, size_factors=sf)
    import ctypes as C
    cnt = (C.c_uint64 * 4)()
    _lib.call("dx_fit_percell_counters", cnt, 1)
    est, sec = timed(run, reps=2)
    _lib.call("dx_fit_percell_counters", cnt, 1)
    passes = {"location_passes_per_tile": cnt[0] / max(cnt[3], 1), "dispersion_passes_per_tile": cnt[1] / max(cnt[3], 1),
              "dispersion_epochs_per_tile": cnt[2] / max(cnt[3], 1), "gene_tiles_per_fit": cnt[3] / 3}
    nit = est.niter
    bytes_iter = 8.0 * shard.nnz + n_cells * (xd.shape[1] * 8 + 8) * n_genes      # design + log sf re-read per gene (L2)
    return {"config": "configs[3]: de.test.wald, %d cells x %d genes (of 30k), 12 spline coefficients, size factors" % (n_cells, n_genes),
            "nnz": shard.nnz, "p_loc": int(xd.shape[1]), "seconds": sec, "genes_per_s": n_genes / sec,
            "device_train_ms": est.device_train_ms,
            "passes": passes,
            "frac_converged": float(np.mean((est.error_codes & 1) != 0)), "mean_niter": float(np.mean(nit)),
            "niter_percentiles_5_25_50_75_95_100": [float(v) for v in np.percentile(nit, [5, 25, 50, 75, 95, 100])],
            "mean_of_tile_max_niter(32 genes)": float(np.mean([nit[i:i + 32].max() for i in range(0, len(nit), 32)])),
            "seconds_per_gene_iteration": sec / float(np.sum(nit)),
            "count_stream_gbs_per_iteration": 8.0 * shard.nnz * float(np.mean(nit)) / sec / 1e9}


def api(n_cells=100_000, n_genes=20_000):
    """configs[1] through the public Python API from a host scipy CSR matrix (what a diffxpy user calls)."""
    import scipy.sparse
    import diffxpy_b200.api as de
    shard, cond, batch = bench.make_shard(torch, dev, seed=2, n_genes=n_genes, n_cells=n_cells)
    csc = scipy.sparse.csc_matrix((shard.vals.cpu().numpy(), shard.rows.cpu().numpy(), shard.indptr.cpu().numpy()),
                                  shape=(n_cells, n_genes))
    csr = csc.tocsr()
    del shard
    torch.cuda.empty_cache()
    sd = pd.DataFrame({"condition": cond.astype(str), "batch": batch.astype(str)})
    names = ["g%d" % i for i in range(n_genes)]
    out = {}
    for label, data in (("csr", csr), ("csc", csc)):
        def run():
            res = de.test.wald(data=data, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                               sample_description=sd, gene_names=names)
            return res.summary()
        summ, sec = timed(run, reps=2)
        out["seconds_" + label] = sec
        out["genes_per_s_" + label] = n_genes / sec
    out.update({"config": "configs[1] via de.test.wald(...).summary() from host scipy sparse (pageable memory), %d x %d" % (n_cells, n_genes),
                "nnz": int(csr.nnz), "n_sig_q05": int((summ["qval"] < 0.05).sum())})
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg5", "cfg3", "cfg4"]
    for w in which:
        out = {"cfg5": cfg5, "cfg3": cfg3, "cfg4": cfg4, "api": api, "ranklog": ranklog}[w]()
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()
