"""CPU, world_size 2 over gloo: the gene-shard plan and the final all-gather of per-gene rows (the only
collective of the path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from diffxpy_b200 import dist as dxdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, g, weights, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=size)
    try:
        w = dxdist.world("auto")
        assert (w.rank, w.size) == (rank, size)
        lo, hi = dxdist.gene_ranges(g, size, weights)[rank]
        # fake per-gene rows: row value encodes the global gene index
        rows = {
            "a_var": torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64),
            "niter": torch.arange(lo, hi, dtype=torch.int32),
            # a [g, 1] column that is the transpose of a [1, g] row (how the estimator hands over b_var): unit-size last
            # dimension with a non-unit stride
            "b_var": torch.arange(lo, hi, dtype=torch.float64)[None, :].t(),
            "fim": torch.arange(lo, hi, dtype=torch.float64)[:, None, None] * torch.ones(1, 2, 2, dtype=torch.float64),
        }
        full = dxdist.gather_genes(w, rows, g)
        # the estimator's call: every rank knows all range lengths, so the count exchange is skipped -- same tables
        known = dxdist.gather_genes(w, rows, g, counts=[b - a for a, b in dxdist.gene_ranges(g, size, weights)])
        for k in rows:
            assert torch.equal(full[k], known[k]) and full[k].dtype == rows[k].dtype
        np.save(os.path.join(outdir, "a_%d.npy" % rank), full["a_var"].numpy())
        np.save(os.path.join(outdir, "n_%d.npy" % rank), full["niter"].numpy())
        np.save(os.path.join(outdir, "f_%d.npy" % rank), full["fim"].numpy())
        assert full["b_var"].shape == (g, 1) and torch.equal(full["b_var"][:, 0], torch.arange(g, dtype=torch.float64))
    finally:
        td.destroy_process_group()


def test_gather_genes_world_size_2(tmp_path):
    g = 37
    weights = np.concatenate([np.full(10, 50), np.full(27, 1)])     # unequal range lengths
    port = _free_port()
    mp.spawn(_worker, args=(2, port, g, weights, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        a = np.load(tmp_path / ("a_%d.npy" % rank))
        n = np.load(tmp_path / ("n_%d.npy" % rank))
        f = np.load(tmp_path / ("f_%d.npy" % rank))
        np.testing.assert_array_equal(a[:, 0], np.arange(g))
        np.testing.assert_array_equal(n, np.arange(g))
        np.testing.assert_array_equal(f[:, 1, 1], np.arange(g))


def test_gene_weights_from_csr_and_csc_agree():
    """Gene cuts by non-zero count for the input diffxpy actually receives (CSR, utils.py:20-33) as for gene-major input."""
    import scipy.sparse
    rng = np.random.default_rng(0)
    dense = rng.poisson(0.3, (50, 23)) * (rng.uniform(size=(50, 23)) < np.linspace(0.05, 0.9, 23)[None, :])
    csr, csc = scipy.sparse.csr_matrix(dense), scipy.sparse.csc_matrix(dense)
    np.testing.assert_array_equal(dxdist.gene_weights(csr), dxdist.gene_weights(csc))
    np.testing.assert_array_equal(dxdist.gene_weights(csr), (dense != 0).sum(axis=0))
    assert dxdist.gene_weights(dense) is None
    assert dxdist.gene_range(23, 1, 3, data=csr) == dxdist.gene_range(23, 1, 3, data=csc)
