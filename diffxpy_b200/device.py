"""
Device-side containers: the gene-major (CSC) count shard that every kernel streams, and thin helpers
that pass torch CUDA tensors to the C ABI as raw pointers.  PyTorch is plumbing here (allocation,
streams, torch.distributed); all arithmetic happens in libdxb200.so.

Layout in HBM (one shard = a contiguous gene range of the cells x genes matrix):
    indptr  int64  [G+1]    start of every gene's run of non-zeros
    rows    int32  [nnz]    cell index, ascending inside a gene
    vals    float32[nnz]    count (exact for integers < 2^24)
i.e. 8 algorithmic bytes per non-zero, the figure SURVEY.md section 8(d) charges against the HBM roofline.
"""
import ctypes as C

import numpy as np
import scipy.sparse
import torch

from . import _lib


class _NothingPacked:
    """Marker of SuffStats.pack_overflow() on a shard without overflow entries: the fit takes a NULL record table."""


NOTHING_PACKED = _NothingPacked()


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None or t is NOTHING_PACKED:
        return None
    assert t.is_contiguous(), "tensor passed to the C ABI must be contiguous"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def cuda_device(device=None):
    _lib.require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


_STAGE = {}     # (device index) -> two pinned staging buffers + their events


_WARNED_F32 = [False]


def _warn_if_rounded_to_float32(a):
    """The shard stores values as float32: exact for counts below 2^24, but float64 normalised / log-transformed values are
    rounded (distinct values may become ties of the rank test; moments change at ~1e-7 relative).  A strided sample of the
    array decides; the warning is logged once per process."""
    if _WARNED_F32[0] or a.dtype != np.float64 or a.size == 0:
        return
    flat = a.reshape(-1)
    sample = flat[::max(1, flat.shape[0] // 65536)]
    if np.any(sample.astype(np.float32).astype(np.float64) != sample):
        _WARNED_F32[0] = True
        import logging
        logging.getLogger("diffxpy").warning(
            "count shard: float64 values that are not representable in float32 are rounded to float32 on the device "
            "(raw counts are unaffected; normalised / log-transformed values keep ~7 significant digits, and values that "
            "differ only beyond that become ties of the rank test)")


def upload(arr, dev, dtype=None):
    """Host ndarray -> device tensor of ``dtype``.  Large pageable arrays go through dx_upload: eight host threads stage
    slices of the array through pinned buffers (converting float64 -> float32 / int64 -> int32 on the way), which keeps
    PCIe busy (~52 GB/s against ~10 for a plain pageable copy) and runs without the GIL."""
    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a)
    dtype = dtype or t.dtype
    if dtype == torch.float32:
        _warn_if_rounded_to_float32(a)
    n = a.size
    kind = None
    if a.ndim == 1 and a.nbytes >= (32 << 20):
        if t.dtype == dtype:
            kind = 0
        elif a.dtype == np.float64 and dtype == torch.float32:
            kind = 1
        elif a.dtype == np.int64 and dtype == torch.int32:
            kind = 2
    if kind is None:
        return t.to(dev).to(dtype)
    out = torch.empty(n, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dx_upload", ptr(out), C.c_void_p(a.ctypes.data), int(a.nbytes if kind == 0 else n), kind, stream_ptr())
    return out


# ---- upload in the background --------------------------------------------------------------------------------------------
# de.test.wald(host matrix, formula, ...) spends tens of milliseconds on the host before the estimator asks for the shard
# (design matrices from the formula, distinct design rows, rank checks) and about as long moving the matrix over PCIe.
# prefetch() starts the second while the first runs: a worker thread uploads (and, for cell-major input, transposes: K8) on
# its own stream; take() hands the shard over, ordered after the worker's stream.
_PREFETCH = {}


def _trace(msg):
    """DXB200_TRACE=1: wall-clock marks of the upload pipeline on stderr (development)."""
    import os
    if os.environ.get("DXB200_TRACE") == "1":
        import sys
        import time
        sys.stderr.write("[dxb200 %.1f ms] %s\n" % (time.perf_counter() * 1e3 % 100000.0, msg))


def prefetch(data, gene_range=None, device=None):
    import threading
    if isinstance(data, CountShard) or not torch.cuda.is_available():
        return
    if hasattr(data, "X") and not isinstance(data, (np.ndarray, torch.Tensor)) and not scipy.sparse.issparse(data):
        return                                   # (AnnData-like: the estimator unwraps it first; uploaded there)
    if not (scipy.sparse.issparse(data) or isinstance(data, np.ndarray)):
        return
    dev = cuda_device(device)
    box = {}

    def work():
        _trace("prefetch worker starts")
        try:
            with torch.cuda.device(dev):
                stream = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(stream):
                    box["shard"] = CountShard.from_host(data, dev, gene_range)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    box["event"] = ev
        except BaseException as exc:          # noqa: BLE001  (re-raised by take())
            box["error"] = exc
        _trace("prefetch worker done (host side)")

    th = threading.Thread(target=work, daemon=True)
    _PREFETCH.clear()                          # at most one matrix in flight
    _PREFETCH[id(data)] = (th, box, gene_range, data)
    th.start()


def take(data, device=None, gene_range=None):
    """The shard of ``data``: the one a prefetch() of the same object and gene range produced, else a fresh upload."""
    ent = _PREFETCH.pop(id(data), None)
    _trace("take() called, prefetched=%s" % (ent is not None))
    if ent is not None and ent[3] is data and ent[2] == gene_range:
        th, box = ent[0], ent[1]
        th.join()
        _trace("take() joined")
        if "error" in box:
            raise box["error"]
        torch.cuda.current_stream().wait_event(box["event"])
        return box["shard"]
    if ent is not None:
        ent[0].join()
    return CountShard.from_host(data, cuda_device(device), gene_range)


class CountShard:
    """Gene-major compressed counts of a gene range on one GPU."""

    def __init__(self, indptr, rows, vals, n_cells, gene_offset=0):
        self.indptr, self.rows, self.vals = indptr, rows, vals
        self.n_cells = int(n_cells)
        self.n_genes = int(indptr.shape[0] - 1)
        self.gene_offset = int(gene_offset)
        self.device = vals.device
        self._nnz = None
        self._order = None
        self._int = None
        self._pieces = None

    @property
    def nnz(self):
        if self._nnz is None:
            self._nnz = int(self.indptr[-1].item())
        return self._nnz

    # ---- what a resident shard knows about itself (computed once, reused by every K_ingest launch) ---------
    @property
    def gene_order(self):
        """Genes sorted by non-zero count, largest first (int32): the order the ingest queue hands them out."""
        if self._order is None:
            counts = self.indptr[1:] - self.indptr[:-1]
            self._order = torch.argsort(counts, descending=True, stable=True).to(torch.int32).contiguous()
        return self._order

    @property
    def integer_counts(self):
        """True when every stored value is a non-negative integer below 2^24 (raw counts)."""
        if self._int is None:
            v = self.vals
            self._int = bool(((v >= 0) & (v < 16777216.0) & (v == torch.floor(v))).all().item()) if v.numel() else True
        return self._int

    @property
    def counts_fit_table(self):
        """True when every stored value is an integer in 1..64: no gene has overflow entries, every count sits in the
        tail-count table (DX_SHARD_NO_OVERFLOW: K_ingest skips the pass that writes overflow lists of cut genes, K_fit the
        launch for genes with overflow entries)."""
        if getattr(self, "_fit64", None) is None:
            v = self.vals
            self._fit64 = bool(self.integer_counts and (v.numel() == 0 or float(v.max().item()) <= 64.0))
        return self._fit64

    INGEST_TILE = 256          # elements per bulk-copy stage of K_ingest: pieces start on multiples of it
    INGEST_WARPS = 20          # resident warps per SM of K_ingest (16 design groups)

    @property
    def work_list(self):
        """(pieces int32[n][4], multi_genes int32[m]) for dx_group_suffstats_pieces: K_ingest's units of work, largest
        first.  A gene is one piece unless it holds more than 60 % of a warp's share of the shard (non-zeros / resident warps),
        in which case it is cut into equal parts of about a quarter of that share (multiples of 256 elements, at least 4096).
        With 20k genes on one GPU no gene is that large and the list is the largest-first gene order; with 2.5k genes (an
        eighth of the matrix) most genes are cut, and the pass stays balanced."""
        if self._pieces is None:
            dev = self.device
            counts = (self.indptr[1:] - self.indptr[:-1])
            g = self.n_genes
            if g == 0:
                self._pieces = (torch.zeros((0, 4), dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
                return self._pieces
            sm = C.c_int(0)
            _lib.call("dx_device_sm_count", C.byref(sm))
            n_warps = max(1, sm.value) * self.INGEST_WARPS
            tile = self.INGEST_TILE
            share = -(-int(self.nnz) // n_warps)                      # non-zeros per warp of an evenly spread pass
            target = max(4096, share // 4)
            target = -(-target // tile) * tile
            cut_above = max((3 * target) // 2, (6 * share) // 10)     # genes beyond 60 % of a warp's share are cut
            n_p = torch.where(counts > cut_above, (counts + target - 1) // target, torch.ones_like(counts))
            part = ((counts + n_p - 1) // n_p + tile - 1) // tile * tile           # elements per part of a cut gene
            n_p = torch.where(n_p > 1, (counts + part - 1) // torch.clamp(part, min=1), n_p)
            n_p = torch.clamp(n_p, min=1)
            gene = torch.repeat_interleave(torch.arange(g, device=dev), n_p)
            first = torch.cumsum(n_p, 0) - n_p
            j = torch.arange(gene.shape[0], device=dev) - first[gene]
            start = j * part[gene]
            length = torch.minimum(part[gene], counts[gene] - start)
            length = torch.where(n_p[gene] > 1, length, counts[gene])
            start = torch.where(n_p[gene] > 1, start, torch.zeros_like(start))
            multi = (n_p[gene] > 1)
            order = torch.argsort(length, descending=True, stable=True)
            pieces = torch.stack([gene, start, length, multi.to(torch.int64)], dim=1)[order].to(torch.int32).contiguous()
            multi_genes = torch.nonzero(n_p > 1, as_tuple=False)[:, 0].to(torch.int32).contiguous()
            self._pieces = (pieces, multi_genes)
        return self._pieces

    def prepare(self):
        """Materialise the cached shard properties (setup time, like the CSR -> CSC conversion)."""
        self.work_list
        self.counts_fit_table
        return self.gene_order, self.integer_counts

    @property
    def algorithmic_bytes(self):
        """Bytes one streaming pass must read: 8 per non-zero + the gene pointers."""
        return 8 * self.nnz + 8 * (self.n_genes + 1)

    # ---- constructors -----------------------------------------------------------------------------
    @staticmethod
    def from_host(data, device=None, gene_range=None):
        """Accepts what diffxpy accepts as ``data`` (utils.py:20-33): dense ndarray, scipy sparse (CSR as
        AnnData stores it, or CSC), an object with ``.X`` (AnnData / Raw) or a CountShard."""
        if isinstance(data, CountShard):
            return data
        dev = cuda_device(device)
        if hasattr(data, "X") and not isinstance(data, (np.ndarray, torch.Tensor)) and not scipy.sparse.issparse(data):
            data = data.X
        if isinstance(data, torch.Tensor):
            return CountShard.from_dense_device(data.to(dev), gene_range)
        if scipy.sparse.issparse(data):
            n_cells = data.shape[0]
            if scipy.sparse.isspmatrix_csc(data):
                csc = data
                if gene_range is not None:
                    lo, hi = gene_range
                    ip = np.asarray(csc.indptr[lo:hi + 1], dtype=np.int64)
                    sl = slice(int(ip[0]), int(ip[-1]))
                    indptr, indices, vals = ip - ip[0], csc.indices[sl], csc.data[sl]
                else:
                    indptr, indices, vals = np.asarray(csc.indptr, dtype=np.int64), csc.indices, csc.data
                if not csc.has_sorted_indices:
                    csc = csc.copy()
                    csc.sort_indices()
                    return CountShard.from_host(csc, dev, gene_range)
                return CountShard(torch.from_numpy(np.ascontiguousarray(indptr)).to(dev),
                                  upload(indices, dev, torch.int32), upload(vals, dev, torch.float32),
                                  n_cells, 0 if gene_range is None else gene_range[0])
            csr = data.tocsr()
            if not csr.has_canonical_format:          # K8 needs every (cell, gene) at most once
                csr = csr.copy()
                csr.sum_duplicates()
            return CountShard.from_csr_device(
                torch.from_numpy(np.asarray(csr.indptr, dtype=np.int64)).to(dev),
                upload(csr.indices, dev, torch.int32), upload(csr.data, dev, torch.float32),
                csr.shape[1], gene_range, sorted_cols=True)          # (canonical format: sorted, no duplicates)
        arr = np.asarray(data)
        if arr.ndim != 2:
            raise ValueError("data must be a 2-d (cells x genes) matrix")
        if gene_range is not None:
            arr = arr[:, gene_range[0]:gene_range[1]]
        _warn_if_rounded_to_float32(arr)
        shard = CountShard.from_dense_device(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(dev))
        shard.gene_offset = 0 if gene_range is None else gene_range[0]
        return shard

    @staticmethod
    def from_dense_device(x, gene_range=None):
        """Dense cells x genes tensor on the device -> shard (zeros dropped)."""
        if gene_range is not None:
            x = x[:, gene_range[0]:gene_range[1]]
        xt = x.t().contiguous().to(torch.float32)
        nz = xt != 0
        counts = nz.sum(dim=1, dtype=torch.int64)
        indptr = torch.zeros(xt.shape[0] + 1, dtype=torch.int64, device=x.device)
        torch.cumsum(counts, 0, out=indptr[1:])
        idx = nz.nonzero(as_tuple=False)
        rows = idx[:, 1].to(torch.int32).contiguous()
        vals = xt[nz].contiguous()
        return CountShard(indptr, rows, vals, x.shape[0], 0 if gene_range is None else gene_range[0])

    @staticmethod
    def from_csr_device(indptr, cols, vals, n_genes, gene_range=None, sorted_cols=False):
        """Cell-major CSR on the device -> gene-major shard of the gene range: K8 (dx_csr_to_csc[_staged]: count per chunk of
        cells, scan, stable placement; cells ascending inside a gene).  A cell must hold a gene at most once; with
        ``sorted_cols`` (column indices ascending inside every cell: canonical CSR) the staged kernels are used."""
        n_cells = indptr.shape[0] - 1
        dev = vals.device
        lo, hi = (0, int(n_genes)) if gene_range is None else (int(gene_range[0]), int(gene_range[1]))
        n_out = hi - lo
        if n_out > 56320:
            return CountShard._from_csr_device_sorted(indptr, cols, vals, n_genes, gene_range)
        nbytes = C.c_int64(0)
        out_ptr = torch.empty(n_out + 1, dtype=torch.int64, device=dev)
        cap = max(int(cols.shape[0]), 1)
        out_rows = torch.empty(cap, dtype=torch.int32, device=dev)
        out_vals = torch.empty(cap, dtype=torch.float32, device=dev)
        import os
        staged = sorted_cols and n_cells < (1 << 27) and os.environ.get("DXB200_K8_STAGED", "1") != "0"          # (test hook)
        if staged:          # append-only segments per gene block first: no write amplification (dx_transpose.cu)
            _lib.call("dx_csr_to_csc_staged_work_bytes", int(n_cells), n_out, cap, C.byref(nbytes))
            work = torch.empty((nbytes.value + 7) // 8, dtype=torch.int64, device=dev)
            _lib.call("dx_csr_to_csc_staged", int(n_cells), cap, ptr(indptr.contiguous()), ptr(cols.contiguous()),
                      ptr(vals.to(torch.float32).contiguous()), lo, n_out, ptr(out_ptr), ptr(out_rows), ptr(out_vals), ptr(work),
                      int(work.numel() * 8), stream_ptr())
        else:
            _lib.call("dx_csr_to_csc_work_bytes", int(n_cells), n_out, C.byref(nbytes))
            work = torch.empty((nbytes.value + 7) // 8, dtype=torch.int64, device=dev)
            _lib.call("dx_csr_to_csc", int(n_cells), ptr(indptr.contiguous()), ptr(cols.contiguous()),
                      ptr(vals.to(torch.float32).contiguous()), lo, n_out, ptr(out_ptr), ptr(out_rows), ptr(out_vals), ptr(work),
                      stream_ptr())
        if gene_range is not None:          # only the entries of the range were written: keep those
            total = int(out_ptr[-1].item())
            out_rows, out_vals = out_rows[:total].clone(), out_vals[:total].clone()
        return CountShard(out_ptr, out_rows, out_vals, n_cells, lo)

    @staticmethod
    def _from_csr_device_sorted(indptr, cols, vals, n_genes, gene_range=None):
        """More genes in the range than K8 has shared-memory counters for: stable sort by gene."""
        n_cells = indptr.shape[0] - 1
        dev = vals.device
        row_of = torch.repeat_interleave(torch.arange(n_cells, device=dev, dtype=torch.int32),
                                         (indptr[1:] - indptr[:-1]))
        if gene_range is not None:
            lo, hi = gene_range
            keep = (cols >= lo) & (cols < hi)
            cols, vals, row_of = cols[keep] - lo, vals[keep], row_of[keep]
            n_out = hi - lo
        else:
            n_out = n_genes
        order = torch.sort(cols.to(torch.int64), stable=True).indices
        counts = torch.bincount(cols.to(torch.int64), minlength=n_out)
        out_ptr = torch.zeros(n_out + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=out_ptr[1:])
        return CountShard(out_ptr, row_of[order].contiguous(), vals[order].to(torch.float32).contiguous(), n_cells,
                          0 if gene_range is None else gene_range[0])

    # ---- kernels ----------------------------------------------------------------------------------
    def group_suffstats(self, cell_group, n_groups):
        """K_ingest: tail counts [G][K][64] uint32 (as int32 storage) + overflow entries."""
        g = self.n_genes
        dev = self.device
        tail = torch.empty((g, n_groups, _lib.DX_HIST_BINS), dtype=torch.int32, device=dev)
        ovf_count = torch.empty(g, dtype=torch.int32, device=dev)
        ovf_val = torch.empty(max(self.vals.shape[0], 1), dtype=torch.float32, device=dev)
        ovf_group = torch.empty(max(self.vals.shape[0], 1), dtype=torch.uint8, device=dev)
        flags = _lib.DX_SHARD_INTEGER_COUNTS if self.integer_counts else 0
        if self.counts_fit_table:
            flags |= _lib.DX_SHARD_NO_OVERFLOW
        if int(torch.bincount(cell_group.to(torch.int64), minlength=256)[:255].max().item()) <= 65535:
            flags |= _lib.DX_GROUPS_FIT_U16
        pieces, multi = self.work_list
        import os
        if g > 0 and (self.rows.data_ptr() | self.vals.data_ptr()) % 16 == 0 and n_groups <= 32 and \
                os.environ.get("DXB200_INGEST_PIECES", "1") != "0":          # (test hook: "0" = the gene-wise entry point)
            _lib.call("dx_group_suffstats_pieces", g, ptr(self.indptr), ptr(self.rows), ptr(self.vals), ptr(cell_group),
                      int(n_groups), ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(pieces), int(pieces.shape[0]),
                      ptr(multi) if multi.shape[0] else None, int(multi.shape[0]), flags, stream_ptr())
        else:
            _lib.call("dx_group_suffstats_ex", g, ptr(self.indptr), ptr(self.rows), ptr(self.vals), ptr(cell_group),
                      int(n_groups), ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(self.gene_order), flags,
                      stream_ptr())
        return SuffStats(self, tail, ovf_count, ovf_val, ovf_group, n_groups)

    def gene_moments(self, inv_sf=None):
        g = self.n_genes
        s1 = torch.empty(g, dtype=torch.float64, device=self.device)
        s2 = torch.empty(g, dtype=torch.float64, device=self.device)
        _lib.call("dx_gene_moments", g, ptr(self.indptr), ptr(self.rows), ptr(self.vals), ptr(inv_sf), ptr(s1), ptr(s2),
                  stream_ptr())
        return s1, s2


class SuffStats:
    """Output of K_ingest for one shard."""

    def __init__(self, shard, tail, ovf_count, ovf_val, ovf_group, n_groups):
        self.shard, self.tail, self.ovf_count, self.ovf_val, self.ovf_group = shard, tail, ovf_count, ovf_val, ovf_group
        self.n_groups = int(n_groups)
        self.ovf_packed = None       # set by pack_overflow(): the overflow arrays are then only valid for the fit

    def pack_overflow(self, min_entries=0):
        """K_pack: rewrite long overflow lists (highly expressed genes) as (value, multiplicity) records, in place.
        Afterwards the statistics feed dx_nb_fit_grouped_packed only (not the rank / t tests)."""
        if self.ovf_packed is None and self.shard.counts_fit_table:
            self.ovf_packed = NOTHING_PACKED          # every count sits in the tail-count table: no list to pack, no launch
        if self.ovf_packed is None:
            packed = torch.empty((self.tail.shape[0], 2), dtype=torch.int32, device=self.tail.device)
            _lib.call("dx_pack_overflow", self.tail.shape[0], self.n_groups, int(min_entries), ptr(self.ovf_count),
                      ptr(self.shard.indptr), ptr(self.ovf_val), ptr(self.ovf_group), ptr(packed), stream_ptr())
            self.ovf_packed = packed
        return self
