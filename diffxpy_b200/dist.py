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


def gene_weights(data):
    """Per-gene non-zero counts of a host matrix (the weights of the gene cuts), or None when counting would cost a pass
    over a matrix that is not sparse.  CSC: the gene pointers; CSR (what AnnData holds): one bincount of the column indices."""
    if data is None or not scipy.sparse.issparse(data):
        return None
    if scipy.sparse.isspmatrix_csc(data):
        return np.diff(data.indptr)
    if scipy.sparse.isspmatrix_csr(data):
        return np.bincount(data.indices, minlength=data.shape[1])
    return None


def gene_range(n_genes, rank, size, data=None):
    weights = gene_weights(data) if size > 1 else None
    return gene_ranges(n_genes, size, weights)[rank]


def gather_genes(w, tensors, n_genes_total, counts=None):
    """All-gather per-gene rows (leading dimension = genes of the local range) into full tables with ONE collective:
    the rows of all tensors are packed side by side into one byte row per gene, padded to the longest gene range, gathered
    (``all_gather_into_tensor``: NCCL over NVLink on GPUs, gloo in the CPU tests) and unpacked.  ``counts``: the lengths of
    all ranks' gene ranges when the caller knows them (the estimator does: every rank computes the same cuts); without it
    one more tiny collective exchanges them."""
    if w.size == 1:
        return tensors
    names = list(tensors)
    any_t = tensors[names[0]]
    g_local, device = int(any_t.shape[0]), any_t.device
    if counts is None:
        local = torch.tensor([g_local], dtype=torch.int64, device=device)
        gathered = torch.empty(w.size, dtype=torch.int64, device=device)
        td.all_gather_into_tensor(gathered, local)
        counts = [int(c) for c in gathered.tolist()]
    assert len(counts) == w.size and counts[w.rank] == g_local and sum(counts) == n_genes_total, "gene ranges disagree"
    gmax = max(max(counts), 1)
    widths, shapes = [], []
    for name in names:
        t = tensors[name]
        assert t.shape[0] == g_local, "every tensor holds one row per local gene"
        shapes.append(tuple(t.shape[1:]))
        row = int(np.prod(t.shape[1:], dtype=np.int64)) * t.element_size()
        widths.append((row + 7) // 8 * 8)          # 8-byte columns keep every field aligned for the typed views
    send = torch.zeros((gmax, sum(widths)), dtype=torch.uint8, device=device)
    col = 0
    for name, wd in zip(names, widths):
        t = tensors[name]
        if g_local:
            ncols = int(np.prod(t.shape[1:], dtype=np.int64))
            rows = torch.empty((g_local, ncols), dtype=t.dtype, device=device)          # (standard strides, whatever the view
            rows.copy_(t.reshape(g_local, ncols))                                       # the caller handed over)
            raw = rows.view(torch.uint8)
            send[:g_local, col:col + raw.shape[1]] = raw
        col += wd
    recv = torch.empty((w.size * gmax, sum(widths)), dtype=torch.uint8, device=device)
    td.all_gather_into_tensor(recv, send)
    recv = recv.view(w.size, gmax, sum(widths))
    out = {}
    col = 0
    for name, wd, shp in zip(names, widths, shapes):
        t = tensors[name]
        nbytes = int(np.prod(shp, dtype=np.int64)) * t.element_size()
        parts = [recv[r, :counts[r], col:col + nbytes] for r in range(w.size)]
        flat = torch.cat(parts, dim=0).contiguous()
        out[name] = flat.view(t.dtype).reshape((n_genes_total,) + shp)
        col += wd
    return out


class PeerExchange:
    """All-gather of float64 result rows over NVLink peer memory (``dx_peer_allgather_f64``): every rank stores its rows
    into the exchange buffer of every peer and waits for the peers' sequence flags -- one small kernel pair instead of
    a library collective (20-40 us on 8 GPUs, and it can sit in a replayed CUDA graph).  torch.distributed is only the plumbing that
    carries the CUDA IPC handles at set-up.  One node, at most 16 ranks, every rank calls ``allgather`` in the same order
    with the same row count."""

    def __init__(self, max_rows_per_rank, slot_rows=None):
        import ctypes as C
        from . import _lib
        if not (td.is_available() and td.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        self.rank, self.size = td.get_rank(), td.get_world_size()
        if self.size > 16:
            raise ValueError("PeerExchange: at most 16 ranks (one node)")
        self._lib, self._C = _lib, C
        # one slot holds the rows of all ranks: size * max_rows_per_rank, or slot_rows when the ranks own ranges of
        # different lengths of one vector of slot_rows entries (gene ranges cut by non-zero count)
        self.slot_bytes = int((slot_rows if slot_rows is not None else self.size * max_rows_per_rank) * 8)
        # Every rank runs the same sequence of collectives whatever fails locally (a rank that raised early would leave
        # the others in a collective it never joins): create, exchange handles, map, agree, then raise everywhere.
        self._local, self._ptrs, err = None, None, None
        handle = C.create_string_buffer(_lib.DX_IPC_HANDLE_BYTES)
        try:
            local = C.c_void_p()
            _lib.call("dx_exchange_create", self.slot_bytes, C.byref(local), handle)
            self._local = local.value
        except Exception as exc:          # noqa: BLE001
            err = exc
        handles = [None] * self.size
        td.all_gather_object(handles, handle.raw if self._local is not None else None)
        ptrs = [None] * self.size
        if err is None and all(h is not None for h in handles):
            try:
                for r in range(self.size):
                    if r == self.rank:
                        ptrs[r] = self._local
                    else:
                        p = C.c_void_p()
                        _lib.call("dx_exchange_open", handles[r], C.byref(p))
                        ptrs[r] = p.value
            except Exception as exc:          # noqa: BLE001
                err = exc
        elif err is None:
            err = RuntimeError("a peer could not create its exchange buffer")
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device="cuda")
        td.all_reduce(ok, op=td.ReduceOp.MIN)
        if int(ok.item()) == 0:
            for r in range(self.size):      # undo what this rank did map
                if r != self.rank and ptrs[r] is not None:
                    try:
                        _lib.call("dx_exchange_close", ptrs[r])
                    except Exception:          # noqa: BLE001
                        pass
            td.barrier()
            if self._local is not None:
                _lib.call("dx_exchange_destroy", self._local)
                self._local = None
            raise RuntimeError("PeerExchange: CUDA IPC set-up failed on at least one rank (%s)" % (err or "a peer failed"))
        self._ptrs = (C.c_void_p * self.size)(*ptrs)
        td.barrier()          # every buffer is mapped everywhere before the first store

    def allgather(self, src, dst):
        """src: float64 CUDA vector of this rank (same length on every rank); dst: float64 [size * len(src)]."""
        from .device import ptr, stream_ptr
        n = int(src.shape[0])
        assert dst.shape[0] == self.size * n and src.dtype == torch.float64 and dst.dtype == torch.float64
        self._lib.call("dx_peer_allgather_f64", n, ptr(src), self._ptrs, self.size, self.rank, self.slot_bytes,
                       ptr(dst), stream_ptr())
        return dst

    @property
    def peer_ptrs(self):
        """ctypes array of the `size` exchange-buffer pointers (own buffer at [rank]) for the C ABI."""
        return self._ptrs

    @property
    def local_ptr(self):
        return self._C.c_void_p(self._local)

    def check(self):
        """Synchronous: raises if a peer's rows did not arrive within the kernel's 5 s limit."""
        w = self._C.c_int32(0)
        self._lib.call("dx_exchange_status", self._local, self._C.byref(w))
        if w.value:
            raise RuntimeError("PeerExchange: rank %d never published its rows" % (w.value - 1))

    def close(self):
        if self._local is None:
            return
        torch.cuda.synchronize()
        td.barrier()          # nobody unmaps a buffer a peer may still write
        for r in range(self.size):
            if r != self.rank:
                self._lib.call("dx_exchange_close", self._ptrs[r])
        td.barrier()
        self._lib.call("dx_exchange_destroy", self._local)
        self._local = None
