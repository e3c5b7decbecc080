#!/usr/bin/env python
"""
bench.py -- headline benchmark of the NB-GLM Wald path (BASELINE.json: "genes/sec, NB-GLM Wald test 100k cells x
20k genes, 1/2/4/8 B200; % HBM roofline").

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA), one rank per GPU under torchrun
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host CPU (oracle port)

Workload = BASELINE.json configs[1]: ONE matrix of 100 000 cells x 20 000 genes, sparse, '~ 1 + condition + batch'
(2 x 8 -> 16 design groups, 9 location coefficients), constant dispersion model, no size factors.  With N GPUs the gene axis
is cut into N contiguous ranges of equal non-zero count (strong scaling; genes are independent, no collective on the data
path); `--scaling weak` gives every GPU its own 20k-gene matrix instead (secondary figure).

One "step" = one full ``de.test.wald`` pass over the shard that is already resident in HBM in gene-major layout, 4 launches:
dx_zero_genes_kernel + K_ingest (the only kernel that streams the count matrix; K_pack, which compresses long overflow lists,
is skipped because the resident shard knows that no count of this workload leaves the tail-count table) -> K_fit (complete batchglm training schedule per gene; its epilogue computes the Wald row and, with N > 1, stores
the p-value straight into every rank's exchange buffer over NVLink) -> BH q-values in one CTA (with N > 1 it first waits for
the exchange flags of all ranks: the one exchange of the path).  With N > 1 the step is replayed as one CUDA graph; at N = 1
it is launched plainly so that the per-kernel CUDA events sit inside the timed region.

Prints ONE JSON line (rank 0).  ``value`` is device-timed whole-job genes/s with inputs resident; ``e2e`` goes through the
public API (de.test.wald(host scipy matrix, formula, sample_description).summary(): uploads, layout conversion, design
handling, gather, statistics and the result table included); ``e2e_cabi`` is the C-ABI host entry point
(dx_wald_grouped_host) from pinned host buffers.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CELLS = 100_000
N_GENES = 20_000
N_BATCH = 8
METRIC = "genes/sec, NB-GLM Wald test 100k cells x 20k genes"
WORKLOAD_STRONG = ("configs[1]: de.test.wald, ONE 100k cells x 20k genes matrix (gene axis split over the GPUs), sparse "
                   "gene-major shard, '~ 1 + condition + batch' (8 batches), constant dispersion")
WORKLOAD_WEAK = ("configs[1] per GPU (weak scaling): de.test.wald, 100k cells x 20k genes on EVERY GPU, sparse gene-major shard, "
                 "'~ 1 + condition + batch' (8 batches), constant dispersion")
WORKLOAD = WORKLOAD_STRONG


def design_for(cond, batch):
    n = cond.shape[0]
    return np.concatenate([np.ones((n, 1)), cond[:, None].astype(np.float64),
                           (batch[:, None] == np.arange(1, N_BATCH)[None, :]).astype(np.float64)], axis=1)


MU_HI = 1.0       # upper end of the log-uniform gene means (tools/kbench.py raises it to probe heavy-tailed shards)


def gene_params(rng, g):
    """SURVEY.md section 8(d) synthetic distribution of configs[1]."""
    log_mu = rng.uniform(np.log(0.005), np.log(MU_HI), g)
    r = rng.uniform(0.5, 4.0, g)
    beta_c = np.where(rng.uniform(size=g) < 0.1, rng.normal(0, 0.5, g), 0.0)
    beta_b = rng.normal(0, 0.2, (N_BATCH, g))
    beta_b[0] = 0.0
    return log_mu, r, beta_c, beta_b


# ------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference path (real batchglm is not installable: no network)
# ------------------------------------------------------------------------------------------------------
def _cpu_fit_chunk(args):
    y, xd = args
    import oracle.nb_glm as onb
    import oracle.stats as ost
    try:        # one BLAS / OpenMP thread per worker process: the workers already cover every core
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    # method_b="newton": the CPU twin of the algorithm the kernels run (same schedule, same scale step).  The scipy-brent
    # variant of the oracle is 1-10x slower per gene (it stalls for 1000 steps on Poisson-like genes), so this is the
    # conservative baseline.
    fit = onb.fit_nb(y, xd, np.ones((y.shape[0], 1)), method_b="newton")
    sd = np.sqrt(np.maximum(fit["fisher_inv"][:, 1, 1], np.nextafter(0, 1)))
    return ost.wald_test(fit["a_var"][1], sd)


def _cpu_warm(_):
    """Start-up of a (spawned) worker outside the timed region: interpreter, numpy / scipy, the oracle modules."""
    import oracle.nb_glm  # noqa: F401
    import oracle.stats  # noqa: F401
    return os.getpid()


def cpu_sample(rng, n_genes_sample, cond, batch):
    log_mu, r, beta_c, beta_b = gene_params(rng, n_genes_sample)
    eta = log_mu[None, :] + cond[:, None] * beta_c[None, :] + beta_b[batch]
    mu = np.exp(eta)
    return rng.poisson(rng.gamma(r[None, :], mu / r[None, :])).astype(np.float64)


def run_cpu_reference(genes_per_step, steps, warmup, cores, seed=1234):
    """Time the oracle port (batchglm numpy schedule: IWLS + safeguarded Newton scale step + reference stats + BH) on
    `genes_per_step` genes of the workload per step, genes spread over `cores` worker processes."""
    import multiprocessing as mp
    import oracle.stats as ost
    rng = np.random.default_rng(seed)
    cond = rng.integers(0, 2, N_CELLS)
    batch = rng.integers(0, N_BATCH, N_CELLS)
    xd = design_for(cond, batch)
    per = max(1, genes_per_step // cores)
    ctx = mp.get_context("spawn")          # clean workers: nothing of the parent's CUDA state is inherited
    times = []
    # The workers are forked from a process that may hold CUDA state.  They never touch it themselves -- but cyclic garbage
    # that holds CUDA tensors (estimators of earlier API calls) would be collected, and its tensors destroyed, by the garbage
    # collector of a CHILD ("CUDA error: initialization error", the worker dies and the pool waits for ever).  Collect it in
    # the parent first and freeze what is alive, so that no child ever frees an object created before the fork.
    import gc
    gc.collect()
    gc.freeze()
    with ctx.Pool(cores) as pool:
        pool.map_async(_cpu_warm, range(4 * cores)).get(timeout=600)
        for it in range(warmup + steps):
            y = cpu_sample(rng, per * cores, cond, batch)
            chunks = [(np.ascontiguousarray(y[:, i * per:(i + 1) * per]), xd) for i in range(cores)]
            t0 = time.perf_counter()
            pv = np.concatenate(pool.map_async(_cpu_fit_chunk, chunks).get(timeout=900))          # (a dead worker raises, not hangs)
            ost.bh_correct(pv)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    gc.unfreeze()
    total = float(np.sum(times))
    return per * cores * len(times) / total, total / len(times), per * cores


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    cores = min(cores, 64)
    genes_per_step = 8 * cores        # 8 genes per worker process and step (a few seconds per step)
    value, sec_per_step, gps = run_cpu_reference(genes_per_step, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_cells": N_CELLS, "genes_total": N_GENES, "genes_per_step": gps,
                   "note": "CPU restatement of batchglm's numpy schedule (oracle/nb_glm.py, dense float64, Newton scale step) "
                           "+ reference stats + BH, one worker process per core; real batchglm is not installable offline"},
        "cpu_baseline": {"value": value, "unit": "genes/s", "cores": cores, "kind": "port",
                         "sample": "%d genes x %d cells per step, %d steps" % (gps, N_CELLS, args.steps)},
        "e2e": {"value": value, "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index, active=True, period_s=0.002):
        super().__init__(daemon=True)
        self.period_s = period_s
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.armed = False          # samples are kept from the start of the timed region on
        self.active = active        # only the rank that prints the line polls: concurrent NVML queries of 8 processes
        #                             serialise in the driver and stall the other ranks' launches
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None or not self.active:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self.stop_flag:
            try:
                clock = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if not self.armed:
                    time.sleep(self.period_s)
                    continue
                self.samples.append(clock)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period_s)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_shard(torch, dev, seed, n_genes=N_GENES, n_cells=N_CELLS, chunk=400):
    """Synthetic NB counts generated on the device, returned as a gene-major CountShard + design info.  Every block of
    `chunk` genes has its own generator seeds, so the matrix does not depend on who generates it: with several GPUs
    every rank builds the same 100k x 20k matrix and keeps its own gene range (strong scaling)."""
    from diffxpy_b200.device import CountShard
    rng = np.random.default_rng(seed)
    cond = rng.integers(0, 2, n_cells)
    batch = rng.integers(0, N_BATCH, n_cells)
    log_mu, r, beta_c, beta_b = gene_params(rng, n_genes)
    gen = torch.Generator(device=dev)
    cond_t = torch.from_numpy(cond.astype(np.float32)).to(dev)
    batch_t = torch.from_numpy(batch).to(dev)
    rows_l, vals_l, counts_l = [], [], []
    for ci, lo in enumerate(range(0, n_genes, chunk)):
        hi = min(lo + chunk, n_genes)
        gen.manual_seed(1_000_003 * seed + 2 * ci)
        torch.cuda.manual_seed(1_000_003 * seed + 2 * ci + 1)   # torch._standard_gamma draws from the default CUDA
        # generator; a different Philox key than `gen`, or the gamma and Poisson draws would be correlated
        lm = torch.from_numpy(log_mu[lo:hi].astype(np.float32)).to(dev)
        bc = torch.from_numpy(beta_c[lo:hi].astype(np.float32)).to(dev)
        bb = torch.from_numpy(beta_b[:, lo:hi].astype(np.float32)).to(dev)
        rr = torch.from_numpy(r[lo:hi].astype(np.float32)).to(dev)
        eta = lm[:, None] + bc[:, None] * cond_t[None, :] + bb.t()[:, batch_t]          # genes x cells
        mu = torch.exp(eta)
        conc = rr[:, None].expand_as(mu).contiguous()
        lam = torch._standard_gamma(conc) * (mu / conc)
        y = torch.poisson(lam, generator=gen)
        nz = y != 0
        counts_l.append(nz.sum(dim=1, dtype=torch.int64))
        idx = nz.nonzero(as_tuple=False)
        rows_l.append(idx[:, 1].to(torch.int32))
        vals_l.append(y[nz].to(torch.float32))
        del eta, mu, conc, lam, y, nz, idx
    counts = torch.cat(counts_l)
    indptr = torch.zeros(n_genes + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    shard = CountShard(indptr, torch.cat(rows_l).contiguous(), torch.cat(vals_l).contiguous(), n_cells)
    return shard, cond, batch


def g_pieces_ok(shard):
    return shard.n_genes > 0 and (shard.rows.data_ptr() | shard.vals.data_ptr()) % 16 == 0


def slice_shard(torch, shard, lo, hi):
    """Gene range [lo, hi) of a resident shard as its own (16-byte aligned) shard."""
    from diffxpy_b200.device import CountShard
    e0, e1 = int(shard.indptr[lo].item()), int(shard.indptr[hi].item())
    return CountShard((shard.indptr[lo:hi + 1] - e0).contiguous(), shard.rows[e0:e1].clone(), shard.vals[e0:e1].clone(),
                      shard.n_cells, gene_offset=lo)


def main_cuda(args):
    import ctypes as C
    import torch
    import torch.distributed as td
    from diffxpy_b200 import _lib
    from diffxpy_b200 import dist as dxdist
    from diffxpy_b200.device import ptr, stream_ptr
    from diffxpy_b200.estimator import group_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    _lib.load()

    strong = args.scaling == "strong"
    n_total_genes, n_cells = (args.genes if strong else args.genes * world), args.cells
    # ---- the workload: ONE 100k x 20k matrix whose gene axis is cut by non-zero count (strong scaling), or
    # one 20k-gene matrix per GPU (--scaling weak, the secondary figure) ----------------------------------------------------
    if strong:
        full, cond, batch = make_shard(torch, dev, seed=1000, n_genes=args.genes, n_cells=n_cells)
        nnz_total = full.nnz
        ranges = dxdist.gene_ranges(args.genes, world, weights=np.diff(full.indptr.cpu().numpy()))
        g_lo, g_hi = ranges[rank]
        shard = slice_shard(torch, full, g_lo, g_hi) if world > 1 else full
        del full
        torch.cuda.empty_cache()
    else:
        shard, cond, batch = make_shard(torch, dev, seed=1000 + rank, n_genes=args.genes, n_cells=n_cells)
        if world > 1:       # the design must be the same on every rank: rank 0's
            cb = [cond, batch] if rank == 0 else [None, None]
            td.broadcast_object_list(cb, src=0)
            # (the counts of the other ranks were drawn under their own design; only the timing matters for this mode)
            cond, batch = cb
        g_lo, g_hi = rank * args.genes, (rank + 1) * args.genes
        nnz_total = None
    gene_order, int_counts = shard.prepare()       # resident-shard metadata (one pass over the values; setup, like the layout itself)
    pieces, multi_genes = shard.work_list          # K_ingest's units of work (whole genes / parts of large genes), same setup
    use_pieces = g_pieces_ok(shard) and os.environ.get("DXB200_BENCH_PIECES", "1") == "1"
    ingest_flags = _lib.DX_SHARD_INTEGER_COUNTS if int_counts else 0
    no_overflow = shard.counts_fit_table          # (same setup pass over the values): every count sits in the tail-count table
    if no_overflow:
        ingest_flags |= _lib.DX_SHARD_NO_OVERFLOW
    nnz = shard.nnz
    g = shard.n_genes
    xd = design_for(cond, batch)
    p, ps = xd.shape[1], 1
    uniq, inv = group_rows(np.concatenate([xd, np.ones((n_cells, 1))], axis=1))
    k = uniq.shape[0]
    xu_loc = np.ascontiguousarray(uniq[:, :p])
    xu_scale = np.ascontiguousarray(uniq[:, p:p + 1])
    n_k = np.bincount(inv, minlength=k).astype(np.float64)
    pinv = np.ascontiguousarray(np.linalg.pinv(xu_loc))
    cell_group = torch.from_numpy(inv.astype(np.uint8)).to(dev)
    if n_k.max() <= 65535:          # what CountShard.group_suffstats checks: the on-chip histograms fit 16-bit bins
        ingest_flags |= _lib.DX_GROUPS_FIT_U16
    d_xl, d_xs = torch.from_numpy(xu_loc).to(dev), torch.from_numpy(xu_scale).to(dev)
    d_nk, d_pinv = torch.from_numpy(n_k).to(dev), torch.from_numpy(pinv).to(dev)
    opts = _lib.DxFitOpts(max_steps=1000, update_b_freq=5, train_loc=1, train_scale=1,
                          init_a=_lib.DX_INIT_CLOSED_FORM, init_b=_lib.DX_INIT_STANDARD, newton_max_iter=100,
                          reserved=_lib.DX_OPT_SCALE_CONST, lltol=1e-10, newton_xtol=1e-8)
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    gq = max(g, 1)
    tail = torch.empty((gq, k, _lib.DX_HIST_BINS), **i32)
    ovf_count = torch.empty(gq, **i32)
    ovf_val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    ovf_group = torch.empty(max(nnz, 1), dtype=torch.uint8, device=dev)
    ovf_packed = torch.empty((gq, 2), **i32)
    a_var, b_var = torch.empty((p, gq), **f64), torch.empty((ps, gq), **f64)
    fim_inv, ll, jac = torch.empty((gq, p, p), **f64), torch.empty(gq, **f64), torch.empty(gq, **f64)
    niter, flags = torch.empty(gq, **i32), torch.empty(gq, **i32)
    rows_out = torch.empty((4, gq), **f64)          # theta, sd, pval, log pval of this rank's genes
    pv_all = torch.empty(n_total_genes, **f64)      # the p-values of all genes (N > 1: gathered through the exchange buffers)
    q_all = torch.empty(n_total_genes, **f64)
    nbytes = C.c_int64(0)
    _lib.call("dx_bh_work_bytes", n_total_genes, C.byref(nbytes))
    bh_work = torch.empty((nbytes.value + 7) // 8, **f64)
    exchange, gather_kind = None, "none"
    if world > 1:
        from diffxpy_b200.dist import PeerExchange
        exchange = PeerExchange(0, slot_rows=n_total_genes)       # raises on every rank together if CUDA IPC is unavailable
        gather_kind = "fit epilogue stores into peer exchange buffers (NVLink), BH kernel waits on the flags"
    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern_events = []

    def step(record=False):
        e = [ev() for _ in range(5)] if record else None
        if record:
            e[0].record()
        if use_pieces:
            _lib.call("dx_group_suffstats_pieces", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cell_group), k,
                      ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(pieces), int(pieces.shape[0]),
                      ptr(multi_genes) if multi_genes.shape[0] else None, int(multi_genes.shape[0]), ingest_flags, stream_ptr())
        else:
            _lib.call("dx_group_suffstats_ex", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cell_group), k,
                      ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(gene_order), ingest_flags, stream_ptr())
        if record:
            e[1].record()
        # what SuffStats.pack_overflow() does in the estimator: K_pack, unless the resident shard knows that no count leaves the
        # tail-count table (this workload: gene means <= 1) -- then there is no list to pack and the fit takes a NULL table
        if not no_overflow:
            _lib.call("dx_pack_overflow", g, k, 0, ptr(ovf_count), ptr(shard.indptr), ptr(ovf_val), ptr(ovf_group),
                      ptr(ovf_packed), stream_ptr())
        if record:
            e[2].record()
        # K_fit + the Wald row of `condition` per gene (+ N > 1: the p-value goes straight into every rank's exchange slot)
        _lib.call("dx_nb_fit_wald_grouped", g, k, p, ps, ptr(d_xl), ptr(d_xs), None, ptr(d_nk), ptr(d_pinv), None, k,
                  ptr(tail), ptr(ovf_count), ptr(shard.indptr), None if no_overflow else ptr(ovf_val),
                  None if no_overflow else ptr(ovf_group), None if no_overflow else ptr(ovf_packed), opts,
                  ptr(a_var), ptr(b_var), ptr(fim_inv), ptr(ll), ptr(jac), ptr(niter), ptr(flags),
                  1, ptr(rows_out[0]), ptr(rows_out[1]), ptr(rows_out[2]), ptr(rows_out[3]),
                  exchange.peer_ptrs if exchange is not None else None, world, rank,
                  exchange.slot_bytes if exchange is not None else 0, g_lo if world > 1 else 0, stream_ptr())
        if record:
            e[3].record()
        if exchange is not None:      # the one exchange of the path: q-values couple all genes
            _lib.call("dx_bh_qvalues_exchange", n_total_genes, exchange.local_ptr, world, exchange.slot_bytes, ptr(pv_all),
                      ptr(q_all), ptr(bh_work), stream_ptr())
        else:
            _lib.call("dx_bh_qvalues", g, ptr(rows_out[2]), ptr(q_all), ptr(bh_work), stream_ptr())
        if record:
            e[4].record()
            kern_events.append(e)
        return q_all

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- before anything is timed: the exchange path must deliver the same vector as the library collective --------------
    step()
    barrier()
    exchange_ok = None
    if world > 1:
        exchange.check()
        counts = [hi - lo for lo, hi in ranges] if strong else [args.genes] * world
        gmax = max(counts)
        pad = torch.full((gmax,), float("nan"), **f64)
        pad[:g] = rows_out[2][:g]
        parts = [torch.empty_like(pad) for _ in range(world)]
        td.all_gather(parts, pad)
        ref_all = torch.cat([parts[r][:counts[r]] for r in range(world)])
        same_p = bool(torch.equal(torch.nan_to_num(ref_all, nan=-1.0), torch.nan_to_num(pv_all, nan=-1.0)))
        from diffxpy_b200.testing.correction import bh_device
        same_q = bool(torch.equal(torch.nan_to_num(bh_device(ref_all), nan=-1.0), torch.nan_to_num(q_all, nan=-1.0)))
        ok_t = torch.tensor([1 if (same_p and same_q) else 0], device=dev)
        td.all_reduce(ok_t, op=td.ReduceOp.MIN)
        exchange_ok = bool(int(ok_t.item()))

    # One CUDA graph per step for N > 1: the launches of a step cost the host 0.1-0.3 ms of Python / driver time, which on
    # an 8-GPU box (8 interpreters sharing the host) shows up as idle gaps on the slowest rank.  At N = 1 the host keeps
    # up, and plain launches let the per-kernel CUDA events sit inside the timed region.  DXB200_BENCH_GRAPH=0 / 1 forces
    # plain launches / the graph.
    graph, launch_kind = None, "eager"
    want_graph = os.environ.get("DXB200_BENCH_GRAPH", "1" if world > 1 else "0") == "1"
    if want_graph:
        try:
            cap_stream = torch.cuda.Stream()
            cap_stream.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream, capture_error_mode="thread_local"):
                step()
            launch_kind = "cuda graph replay (one graph per step)"
        except Exception as exc:          # noqa: BLE001
            graph = None
            if rank == 0:
                print("bench: graph capture failed (%s), timing plain launches" % str(exc).splitlines()[0], file=sys.stderr)
        if world > 1:
            ok = torch.tensor([1 if graph is not None else 0], device=dev)
            td.all_reduce(ok, op=td.ReduceOp.MIN)          # all ranks launch the same way
            if int(ok.item()) == 0:
                graph, launch_kind = None, "eager"
    run_step = (lambda: graph.replay()) if graph is not None else (lambda: step(record=False))
    timed_step = (lambda: graph.replay()) if graph is not None else (lambda: step(record=True))

    # NVML is initialised and the polling thread started BEFORE the barrier (nvmlInit takes milliseconds: done after it,
    # rank 0 would enter the timed region late and every other rank would wait for it in the first exchange)
    # (several processes per node: NVML queries take a driver-wide lock that delays the other ranks' launches -> poll less often)
    sampler = ClockSampler(local_rank, active=(rank == 0), period_s=0.002 if world == 1 else 0.005)
    sampler.start()
    for _ in range(args.warmup):
        run_step()
    barrier()
    t_start, t_stop = ev(), ev()
    t_start.record()
    sampler.armed = True
    for _ in range(args.steps):
        timed_step()
    t_stop.record()
    barrier()
    elapsed_ms = t_start.elapsed_time(t_stop)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    if graph is not None:
        # per-kernel times: events cannot sit inside a replayed graph, so a few more steps with plain launches follow
        for _ in range(min(args.steps, 5)):
            step(record=True)
        barrier()
    if exchange is not None:
        exchange.check()
    ingest_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in kern_events]))
    pack_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in kern_events]))
    fit_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in kern_events]))
    bh_ms = float(np.mean([e[3].elapsed_time(e[4]) for e in kern_events]))       # N > 1: includes waiting for the slowest rank
    ms_per_step = elapsed_ms / args.steps
    value = n_total_genes * args.steps / (elapsed_ms * 1e-3)
    per_rank = None
    if world > 1:          # every rank's own view of the step (the printed kernel times are rank 0's)
        mine = {"genes": g, "nnz": nnz, "ingest": round(ingest_ms, 4), "pack": round(pack_ms, 4), "fit_wald": round(fit_ms, 4),
                "bh_incl_wait": round(bh_ms, 4)}
        per_rank = [None] * world
        td.all_gather_object(per_rank, mine)
    nnz_all = nnz
    if world > 1:
        tt = torch.tensor([nnz], dtype=torch.int64, device=dev)
        td.all_reduce(tt)
        nnz_all = int(tt.item())
    mean_iter = float(niter[:g].to(torch.float64).mean().item()) if g else 0.0
    max_iter = int(niter[:g].max().item()) if g else 0
    conv = float((flags[:g] & 1).to(torch.float64).mean().item()) if g else 1.0
    niter_dev = niter[:g].cpu().numpy().copy()

    # ---- e2e (1): the public API, host matrices in, result table out ----------------------------------------------------
    # de.test.wald(data=<scipy matrix on the host>, formula_loc=..., sample_description=...).summary(): design matrices
    # from the formula, design grouping, upload (pageable memory through pinned staging), layout conversion for CSR input,
    # ingest + fit, gather over the ranks, Wald statistics, BH and the summary DataFrame -- everything a user pays.
    # N > 1: gene-major input so that every rank uploads its own gene range only.
    import pandas as pd
    import scipy.sparse
    import diffxpy_b200.api as de
    full_host = None
    if strong:
        # the complete matrix on the host (every rank: the API shards it itself)
        if world > 1:
            del shard
            torch.cuda.empty_cache()
            full, _, _ = make_shard(torch, dev, seed=1000, n_genes=args.genes, n_cells=n_cells)
        else:
            full = shard
        h_indptr, h_rows, h_vals = full.indptr.cpu().numpy(), full.rows.cpu().numpy(), full.vals.cpu().numpy()
        csc = scipy.sparse.csc_matrix((h_vals, h_rows, h_indptr), shape=(n_cells, args.genes))
        full_host = csc if world > 1 else csc.tocsr()          # CSR (cells x genes, what AnnData holds) on one GPU
        if world > 1:
            del full
            torch.cuda.empty_cache()
    # Two warm-up calls, then `api_steps` timed calls; the line carries every call's time next to the mean: the host side of a
    # call (pageable-memory copies into the staging buffers at 35-44 GB/s, design handling) depends on how busy the box's host
    # is -- on a box that has just come up, and right after the fork pool of the CPU baseline, calls were seen 2-5x slower,
    # which is why this leg runs before the CPU baseline.
    api_s = api_steps = None
    api_same = None
    api_calls_ms = None
    sample_description = pd.DataFrame({"condition": cond.astype(str), "batch": batch.astype(str)})
    gene_names = np.array(["g%d" % i for i in range(args.genes)])
    if full_host is not None and not args.no_api_e2e:
        def api_step():
            res = de.test.wald(data=full_host, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                               sample_description=sample_description, gene_names=gene_names, noise_model="nb")
            return res, res.summary()
        for _ in range(2):
            res, summ = api_step()
        barrier()
        api_steps = max(1, min(args.steps, 5))
        api_calls_ms = []
        t0 = time.perf_counter()
        for _ in range(api_steps):
            tc = time.perf_counter()
            res, summ = api_step()
            api_calls_ms.append(round((time.perf_counter() - tc) * 1e3, 2))
        barrier()
        api_s = time.perf_counter() - t0
        t = torch.tensor([api_s], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        api_s = float(t.item())
        api_same = bool(np.array_equal(np.asarray(res.model_estim.niter)[g_lo:g_hi], niter_dev)) and summ.shape[0] == args.genes
        del res, summ

    # ---- e2e (2): the C-ABI host entry point from pinned host buffers (H2D + compute + D2H per step), N = 1 ----------------
    cabi = None
    if world == 1:
        pin = lambda tns: tns.cpu().pin_memory()
        hp_indptr, hp_rows, hp_vals, hp_grp = pin(shard.indptr), pin(shard.rows), pin(shard.vals), pin(cell_group)
        h_out = {name: torch.empty(shape, dtype=dt).pin_memory() for name, shape, dt in [
            ("a", (p, g), torch.float64), ("b", (ps, g), torch.float64), ("fi", (g, p, p), torch.float64),
            ("ll", (g,), torch.float64), ("jac", (g,), torch.float64), ("ni", (g,), torch.int32), ("fl", (g,), torch.int32),
            ("pv", (g,), torch.float64), ("lp", (g,), torch.float64), ("sd", (g,), torch.float64), ("q", (g,), torch.float64)]}
        hp = lambda tns: C.c_void_p(tns.data_ptr())
        npp = lambda arr: arr.ctypes.data_as(C.c_void_p)

        def e2e_step():
            _lib.call("dx_wald_grouped_host", n_cells, g, hp(hp_indptr), hp(hp_rows), hp(hp_vals), hp(hp_grp), k, p, ps,
                      npp(xu_loc), npp(xu_scale), None, npp(n_k), npp(pinv), None, k, C.byref(opts), 1,
                      hp(h_out["a"]), hp(h_out["b"]), hp(h_out["fi"]), hp(h_out["ll"]), hp(h_out["jac"]), hp(h_out["ni"]),
                      hp(h_out["fl"]), hp(h_out["pv"]), hp(h_out["lp"]), hp(h_out["sd"]), hp(h_out["q"]))

        e2e_steps = max(1, min(args.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        h2d = 8 * nnz + 8 * (g + 1) + n_cells + 8 * (k * (p + ps + 1) + p * k)
        d2h = g * (8 * (p + ps + p * p + 2 + 4) + 8)
        same = bool(np.array_equal(h_out["ni"].numpy(), niter_dev)) and \
            bool(np.array_equal(np.nan_to_num(h_out["q"].numpy(), nan=-1.0), np.nan_to_num(q_all.cpu().numpy(), nan=-1.0)))
        cabi = {"value": g * e2e_steps / e2e_s, "unit": "genes/s", "ms_per_call": e2e_s / e2e_steps * 1e3,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "api": "dx_wald_grouped_host (C ABI, pinned host CSC in; per-gene rows incl. q-values out)",
                "matches_device_path": same}
        _lib.call("dx_host_release")

    # ---- roofline --------------------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured" if peaks else "fallback 6650 GB/s, of fallback"
    # SURVEY 8(d) algorithmic bytes of one pass: 8 B per non-zero + the gene pointers + the cell -> design-group table; the
    # tail-count table the kernel writes (4 G K 64) is its own output format and is NOT counted
    ingest_bytes = 8 * nnz + 8 * (g + 1) + n_cells
    step_bytes_all = 8 * nnz_all + 8 * (n_total_genes + world) + n_cells * world
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    sm_count = C.c_int(0)
    _lib.call("dx_device_sm_count", C.byref(sm_count))
    clocks = sampler.summary()
    kern = {"dx_ingest_tma_kernel": {"ms": ingest_ms, "algorithmic_bytes": ingest_bytes,
                                     "gbs": ingest_bytes / (ingest_ms * 1e-3) / 1e9, "frac_of_hbm_peak": ingest_bytes / (ingest_ms * 1e-3) / 1e9 / peak},
            ("dx_pack_overflow_kernel" if not no_overflow else "dx_pack_overflow_kernel (skipped: no count leaves the table)"): {"ms": pack_ms},
            "dx_fit_lane_kernel(+wald epilogue)": {"ms": fit_ms},
            "dx_bh_cluster_kernel": {"ms": bh_ms}}
    fit_prof = prof.get("dx_fit_lane_kernel_profile")
    if isinstance(fit_prof, dict) and fit_prof.get("warp_instructions") and world == 1 and g == fit_prof.get("genes"):
        # K_fit is bounded by instruction issue, not by HBM (it reads 4 KB of statistics per gene): achieved = warp
        # instructions of one launch (ncu, profiles/) / live duration; peak = SMs x 4 schedulers x SM clock
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        issue_peak = sm_count.value * 4 * sm_mhz * 1e6
        ach = fit_prof["warp_instructions"] / (fit_ms * 1e-3)
        kern["dx_fit_lane_kernel(+wald epilogue)"]["roofline"] = {
            "bound": "issue", "achieved": ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s",
            "frac": ach / issue_peak, "fp64_pipe_pct": fit_prof.get("fp64_pipe_pct"),
            "source": "profiles/traffic.json (ncu smsp__inst_executed.sum of one launch) / live CUDA-event time"}
    dom = "dx_ingest_tma_kernel"       # the kernel of the step that streams HBM (K <= 32 groups, aligned shard: TMA variant)
    achieved = ingest_bytes / (ingest_ms * 1e-3) / 1e9
    traffic = prof.get(dom) if (world == 1 and strong and args.genes == N_GENES and n_cells == N_CELLS) else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "stored: ncu --set full capture of the same command (profiles/traffic.json), not measured in this run",
                "peak_source": peak_source,
                "share_of_step": ingest_ms / ms_per_step,
                "algorithmic_bytes": "8 nnz + 8 (G + 1) + n_cells of this rank (SURVEY 8d inputs only)",
                "whole_step": {"achieved": step_bytes_all / (ms_per_step * 1e-3) / 1e9, "peak": peak * world,
                               "frac": step_bytes_all / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                               "note": "all ranks' input bytes / max-over-ranks step time / (N x peak): the aggregate figure of "
                                       "BASELINE.json's metric (every IRLS / Newton iteration runs from on-chip statistics, so "
                                       "one pass over the matrix is all the HBM traffic of the whole fit)"},
                "note": "K_fit (%.3f ms) runs all IRLS/Newton iterations from on-chip statistics and is issue/latency bound" % fit_ms}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, 64)
        v, sec, gps = run_cpu_reference(16 * cores, steps=1, warmup=0, cores=cores)
        cpu = {"value": v, "unit": "genes/s", "cores": cores, "kind": "port",
               "sample": "%d genes x %d cells, 1 pass (oracle port of batchglm's numpy schedule, Newton scale step, "
                         "one process per core)" % (gps, N_CELLS)}

    # own kernels per step: ingest, pack, fit (+ Wald epilogue, + peer stores), BH (one cluster kernel up to 32768 p-values,
    # else wait/copy + 7 kernels)
    # (+ dx_zero_genes_kernel when the work list cuts large genes into parts)
    launches_per_step = (4 if n_total_genes <= 32768 else (3 + (1 if world > 1 else 0) + 7)) + \
        (1 if (use_pieces and int(multi_genes.shape[0]) > 0) else 0) - (1 if no_overflow else 0)
    if rank == 0:
        if api_s is not None:
            api_bytes = 8 * nnz_all // world + 8 * (n_cells + 1) + n_cells * (p + 1) * 8
            e2e = {"value": args.genes * api_steps / api_s, "unit": "genes/s", "ms_per_call": api_s / api_steps * 1e3,
                   "h2d_bytes_per_step": int(api_bytes), "d2h_bytes_per_step": int((g_hi - g_lo) * (8 * (p + ps + p * p + 2) + 8)),
                   "steps": api_steps,
                   "api": "de.test.wald(data=scipy %s on the host, formula_loc='~ 1 + condition + batch', factor_loc_totest="
                          "'condition', sample_description=DataFrame).summary()" % ("CSR" if world == 1 else "CSC"),
                   "host_memory": "pageable", "ms_calls": api_calls_ms, "ms_median": float(np.median(api_calls_ms)),
                   "warmup_calls": 2, "matches_device_path": api_same}
        elif cabi is not None:
            e2e = cabi
        else:
            e2e = None
        line = {
            "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_STRONG if strong else WORKLOAD_WEAK, "n_cells": n_cells, "genes_total": n_total_genes,
                       "genes_this_rank": g, "nnz_total": nnz_all, "nnz_this_rank": nnz,
                       "partition": "contiguous gene ranges with equal non-zero counts (prefix sum over the gene pointers)" if strong
                       else "one matrix per GPU",
                       "density": nnz_all / (n_cells * n_total_genes), "design_groups": k, "p_loc": p, "p_scale": ps,
                       "l2": "inputs (%.2f GB per pass and GPU) %s L2; no explicit flush" %
                             (8 * nnz / 1e9, "larger than" if 8 * nnz > 126e6 * 2 else "comparable to"),
                       "frac_converged": conv, "mean_niter": mean_iter, "max_niter": max_iter,
                       "launch": launch_kind, "exchange": gather_kind, "exchange_matches_nccl": exchange_ok,
                       "setup_outside_timed_region": "gene-major layout, CountShard.prepare() (one pass over the values: integer "
                                                     "check + largest-first gene order), design grouping",
                       "note": "parity oracle = CPU restatement of batchglm's numpy schedule (oracle/nb_glm.py); real batchglm is "
                               "not installable offline, parity against it is unpinned"},
            "clocks": clocks,
            "e2e": e2e,
            "e2e_cabi": cabi,
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "kernels": kern,
            "per_rank_ms": per_rank,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        if exchange is not None:
            exchange.close()
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="wald", choices=["wald", "lrt", "spline", "rank"],
                    help="wald (default): the headline configs[1]; lrt / spline / rank: configs[2] / [3] / [4] (bench_workloads.py)")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--genes", type=int, default=None, help="genes of the problem (strong) / per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): ONE 100k x 20k problem split over the GPUs by non-zero count; weak: 20k genes per GPU")
    ap.add_argument("--no-api-e2e", action="store_true")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload != "wald" and args.impl == "cuda":
        import bench_workloads
        d_steps, d_warm = bench_workloads.DEFAULT_STEPS[args.workload]
        args.steps = d_steps if args.steps is None else args.steps
        args.warmup = max(d_warm if args.warmup is None else args.warmup, 1)
        return bench_workloads.main(args, sys.modules[__name__])
    args.steps = 20 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else args.warmup
    args.genes = N_GENES if args.genes is None else args.genes
    args.cells = N_CELLS if args.cells is None else args.cells
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return main_reference(args)
    return main_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
