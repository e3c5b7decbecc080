"""
Multi-GPU plumbing: genes are independent, so the gene axis is cut into one
contiguous range per rank (equal non-zero counts when the host matrix is gene-major, equal gene counts
otherwise), every rank fits its own range with no collective on the data path, and ONE all-gather of the
per-gene result rows (NCCL over NVLink on GPUs, gloo in the CPU tests) makes the full table available on
every rank.  One process per GPU, launched by torchrun; with a single process everything is a no-op.
"""
from collections import namedtuple

import numpy as np
import scipy.sparse
import torch
import torch.distributed as td

World = namedtuple("World", ["rank", "size"])


def world(mode="auto"):
    """The (rank, size) the estimator shards over. ``mode``: 'auto' (use torch.distributed if it is
    initialised), False/None/'off' (single process)."""
    if mode in (False, None, "off"):
        return World(0, 1)
    if td.is_available() and td.is_initialized():
        return World(td.get_rank(), td.get_world_size())
    return World(0, 1)


def gene_ranges(n_genes, size, weights=None):
    """Contiguous [lo, hi) gene ranges, one per rank.  With ``weights`` (per-gene non-zero counts) the cut
    points equalise the weight prefix sum, otherwise the gene count."""
    if size <= 1:
        return [(0, int(n_genes))]
    if weights is None:
        cuts = [int(round(i * n_genes / size)) for i in range(size + 1)]
    else:
        w = np.asarray(weights, dtype=np.float64) + 1.0     # +1: empty genes still cost a fit
        cs = np.concatenate([[0.0], np.cumsum(w)])
        targets = cs[-1] * np.arange(size + 1) / size
        cuts = [int(c) for c in np.searchsorted(cs, targets, side="left")]
        cuts[0], cuts[-1] = 0, int(n_genes)
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(size)]


def gene_range(n_genes, rank, size, data=None):
    weights = None
    if size > 1 and data is not None and scipy.sparse.issparse(data) and scipy.sparse.isspmatrix_csc(data):
        weights = np.diff(data.indptr)
    return gene_ranges(n_genes, size, weights)[rank]


def gather_genes(w, tensors, n_genes_total):
    """All-gather per-gene rows (leading dimension = genes of the local range) into full tables.
    Ranges may have different lengths: rows are padded to the longest range for the collective."""
    if w.size == 1:
        return tensors
    out = {}
    any_t = next(iter(tensors.values()))
    local = torch.tensor([any_t.shape[0]], dtype=torch.int64, device=any_t.device)
    counts = [torch.zeros_like(local) for _ in range(w.size)]
    td.all_gather(counts, local)
    counts = [int(c.item()) for c in counts]
    gmax = max(counts)
    for name, t in tensors.items():
        pad = torch.zeros((gmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(w.size)]
        td.all_gather(parts, pad.contiguous())
        out[name] = torch.cat([parts[r][:counts[r]] for r in range(w.size)], dim=0)
        assert out[name].shape[0] == n_genes_total, "gathered gene count mismatch"
    return out
