#!/usr/bin/env python
"""
bench.py -- headline benchmark of the NB-GLM Wald path (BASELINE.json: "genes/sec, NB-GLM Wald test 100k cells x
20k genes, 1/2/4/8 B200; % HBM roofline").

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA), one rank per GPU under torchrun
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host CPU (oracle port)

One "step" = one full ``de.test.wald`` pass over one shard of synthetic counts that is already resident in HBM
as a gene-major CSC shard: K_ingest (the only kernel that streams the count matrix) -> K_pack (compresses long
overflow lists; nothing to do on this workload) -> K_fit (complete batchglm training schedule per gene) -> Wald
statistics -> (N > 1: all-gather of the per-gene p-values over NVLink peer memory) -> BH q-values.  With N > 1 the step
is replayed as one CUDA graph; at N = 1 it is launched plainly so that the per-kernel CUDA events sit inside the timed
region.
Workload = BASELINE.json configs[1]: 100 000 cells x 20 000 genes PER GPU (weak scaling: genes are independent,
every rank owns its own 20k-gene shard; no collective on the data path), sparse, '~ 1 + condition + batch'
(2 x 8 -> 16 design groups, 9 location coefficients), constant dispersion model, no size factors.

Prints ONE JSON line (rank 0).  ``value`` is device-timed whole-job genes/s with inputs resident; ``e2e`` goes
through the C-ABI host entry point (dx_wald_grouped_host) from pinned host buffers, copies included.
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
WORKLOAD = ("configs[1]: de.test.wald, 100k cells x 20k genes per GPU, sparse CSC shard, "
            "'~ 1 + condition + batch' (8 batches), constant dispersion")


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
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(warmup + steps):
            y = cpu_sample(rng, per * cores, cond, batch)
            chunks = [(np.ascontiguousarray(y[:, i * per:(i + 1) * per]), xd) for i in range(cores)]
            t0 = time.perf_counter()
            pv = np.concatenate(pool.map(_cpu_fit_chunk, chunks))
            ost.bh_correct(pv)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_cells": N_CELLS, "genes_per_step": gps,
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
    """Synthetic NB counts generated on the device, returned as a gene-major CountShard + design info."""
    from diffxpy_b200.device import CountShard
    rng = np.random.default_rng(seed)
    cond = rng.integers(0, 2, n_cells)
    batch = rng.integers(0, N_BATCH, n_cells)
    log_mu, r, beta_c, beta_b = gene_params(rng, n_genes)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    torch.cuda.manual_seed(7919 * seed + 17)   # torch._standard_gamma draws from the default CUDA generator; a
    # different Philox key than `gen`, or the gamma and Poisson draws would be correlated
    cond_t = torch.from_numpy(cond.astype(np.float32)).to(dev)
    batch_t = torch.from_numpy(batch).to(dev)
    rows_l, vals_l, counts_l = [], [], []
    for lo in range(0, n_genes, chunk):
        hi = min(lo + chunk, n_genes)
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


def main_cuda(args):
    import torch
    import torch.distributed as td
    from diffxpy_b200 import _lib
    from diffxpy_b200.device import ptr, stream_ptr
    from diffxpy_b200.estimator import group_rows
    from diffxpy_b200.testing.correction import bh_device

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

    n_genes, n_cells = args.genes, args.cells
    shard, cond, batch = make_shard(torch, dev, seed=1000 + rank, n_genes=n_genes, n_cells=n_cells)
    gene_order, int_counts = shard.prepare()       # resident-shard metadata (setup, like the CSC layout itself)
    ingest_flags = _lib.DX_SHARD_INTEGER_COUNTS if int_counts else 0
    nnz = shard.nnz
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
                          reserved=0, lltol=1e-10, newton_xtol=1e-8)
    g = n_genes
    f64 = dict(dtype=torch.float64, device=dev)
    tail = torch.empty((g, k, _lib.DX_HIST_BINS), dtype=torch.int32, device=dev)
    ovf_count = torch.empty(g, dtype=torch.int32, device=dev)
    ovf_val = torch.empty(nnz, dtype=torch.float32, device=dev)
    ovf_group = torch.empty(nnz, dtype=torch.uint8, device=dev)
    ovf_packed = torch.empty((g, 2), dtype=torch.int32, device=dev)
    a_var, b_var = torch.empty((p, g), **f64), torch.empty((ps, g), **f64)
    fim_inv, ll, jac = torch.empty((g, p, p), **f64), torch.empty(g, **f64), torch.empty(g, **f64)
    niter, flags = torch.empty(g, dtype=torch.int32, device=dev), torch.empty(g, dtype=torch.int32, device=dev)
    rows_out = torch.empty((4, g), **f64)          # theta, sd, pval, log pval
    pv_all = torch.empty(world * g, **f64) if world > 1 else None
    exchange, gather_kind = None, "none"
    if world > 1:
        gather_kind = "nccl"
        if os.environ.get("DXB200_GATHER", "peer") == "peer":
            try:          # the product's own gather over NVLink peer memory; NCCL when CUDA IPC is not available
                from diffxpy_b200.dist import PeerExchange
                exchange = PeerExchange(g)
                gather_kind = "peer"
            except Exception as exc:          # noqa: BLE001
                if rank == 0:
                    print("bench: peer exchange unavailable (%s), using NCCL" % exc, file=sys.stderr)
                exchange = None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern_events = []

    def step(record=False):
        e = [ev() for _ in range(7)] if record else None
        if record:
            e[0].record()
        _lib.call("dx_group_suffstats_ex", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cell_group), k,
                  ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(gene_order), ingest_flags, stream_ptr())
        if record:
            e[1].record()
        # what SuffStats.pack_overflow() does in the estimator: nothing to pack on this workload (gene means <= 1)
        _lib.call("dx_pack_overflow", g, k, 0, ptr(ovf_count), ptr(shard.indptr), ptr(ovf_val), ptr(ovf_group),
                  ptr(ovf_packed), stream_ptr())
        if record:
            e[4].record()
        _lib.call("dx_nb_fit_grouped_packed", g, k, p, ps, ptr(d_xl), ptr(d_xs), None, ptr(d_nk), ptr(d_pinv), None, k,
                  ptr(tail), ptr(ovf_count), ptr(shard.indptr), ptr(ovf_val), ptr(ovf_group), ptr(ovf_packed), opts,
                  ptr(a_var), ptr(b_var), ptr(fim_inv), ptr(ll), ptr(jac), ptr(niter), ptr(flags), stream_ptr())
        if record:
            e[2].record()
        _lib.call("dx_wald_from_fit", g, p, 1, ptr(a_var), ptr(fim_inv), ptr(rows_out[0]), ptr(rows_out[1]),
                  ptr(rows_out[2]), ptr(rows_out[3]), stream_ptr())
        if record:
            e[3].record()
            kern_events.append(e)
        pv = rows_out[2]
        if world > 1:      # the one collective of the path: the per-gene p-values of every rank (q-values couple all genes)
            if exchange is not None:
                exchange.allgather(pv, pv_all)
            else:
                td.all_gather_into_tensor(pv_all, pv)
            pv = pv_all
        if record:
            e[5].record()
        q = bh_device(pv)
        if record:
            e[6].record()
        return q

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # One CUDA graph per step: the 11-12 launches of a step cost the host 0.1-0.3 ms of Python / driver time, which on an
    # 8-GPU box (8 interpreters sharing the host) shows up as idle gaps on the slowest rank (1.32 -> see profiles/).  Used
    # for N > 1 only: at N = 1 the host keeps up, and plain launches let the per-kernel CUDA events sit inside the timed
    # region.  DXB200_BENCH_GRAPH=0 / 1 forces plain launches / the graph.
    graph, launch_kind = None, "eager"
    step()
    barrier()
    want_graph = os.environ.get("DXB200_BENCH_GRAPH", "1" if world > 1 else "0") == "1"
    if want_graph:
        try:
            cap_stream = torch.cuda.Stream()
            cap_stream.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream, capture_error_mode="thread_local"):
                q = step()
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
    # rank 0 would enter the timed region late and every other rank would wait for it in the first gather)
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
            q = step(record=True)
        barrier()
    ingest_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in kern_events]))
    pack_ms = float(np.mean([e[1].elapsed_time(e[4]) for e in kern_events]))
    fit_ms = float(np.mean([e[4].elapsed_time(e[2]) for e in kern_events]))
    wald_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in kern_events]))
    gather_ms = float(np.mean([e[3].elapsed_time(e[5]) for e in kern_events]))       # includes waiting for the slowest rank
    bh_ms = float(np.mean([e[5].elapsed_time(e[6]) for e in kern_events]))
    value = world * g * args.steps / (elapsed_ms * 1e-3)
    per_rank = None
    if world > 1:          # every rank's own view of the step (the printed kernel times are rank 0's)
        mine = {"ingest": round(ingest_ms, 4), "pack": round(pack_ms, 4), "fit": round(fit_ms, 4), "wald": round(wald_ms, 4),
                "gather_incl_wait": round(gather_ms, 4), "bh": round(bh_ms, 4)}
        per_rank = [None] * world
        td.all_gather_object(per_rank, mine)

    # ---- e2e: C-ABI host entry point from pinned host buffers (H2D + compute + D2H per step) -------------
    pin = lambda tns: tns.cpu().pin_memory()
    h_indptr, h_rows, h_vals, h_grp = pin(shard.indptr), pin(shard.rows), pin(shard.vals), pin(cell_group)
    h_out = {name: torch.empty(shape, dtype=dt).pin_memory() for name, shape, dt in [
        ("a", (p, g), torch.float64), ("b", (ps, g), torch.float64), ("fi", (g, p, p), torch.float64),
        ("ll", (g,), torch.float64), ("jac", (g,), torch.float64), ("ni", (g,), torch.int32), ("fl", (g,), torch.int32),
        ("pv", (g,), torch.float64), ("lp", (g,), torch.float64), ("sd", (g,), torch.float64)]}
    import ctypes as C
    hp = lambda tns: C.c_void_p(tns.data_ptr())
    npp = lambda arr: arr.ctypes.data_as(C.c_void_p)

    def e2e_step():
        _lib.call("dx_wald_grouped_host", n_cells, g, hp(h_indptr), hp(h_rows), hp(h_vals), hp(h_grp), k, p, ps,
                  npp(xu_loc), npp(xu_scale), None, npp(n_k), npp(pinv), None, k, C.byref(opts), 1,
                  hp(h_out["a"]), hp(h_out["b"]), hp(h_out["fi"]), hp(h_out["ll"]), hp(h_out["jac"]), hp(h_out["ni"]),
                  hp(h_out["fl"]), hp(h_out["pv"]), hp(h_out["lp"]), hp(h_out["sd"]))

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * g * e2e_steps / e2e_s
    h2d = 8 * nnz + 8 * (g + 1) + n_cells + 8 * (k * (p + ps + 1) + p * k)
    d2h = g * (8 * (p + ps + p * p + 2 + 3) + 8)
    same = bool(np.array_equal(h_out["ni"].numpy(), niter.cpu().numpy()))

    # ---- roofline of the kernel that dominates the step -------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ingest_bytes = 8 * nnz + 8 * (g + 1) + n_cells + 4 * g * k * _lib.DX_HIST_BINS + 4 * g
    fit_bytes = 4 * g * k * _lib.DX_HIST_BINS + 8 * g * (p + ps + p * p + 2) + 8 * g + 8 * (g + 1)
    kern = {"dx_ingest_tma_kernel": {"ms": ingest_ms, "algorithmic_bytes": ingest_bytes},
            "dx_pack_overflow_kernel": {"ms": pack_ms},
            "dx_fit_grouped_kernel": {"ms": fit_ms, "algorithmic_bytes": fit_bytes},
            "dx_wald_from_fit_kernel": {"ms": wald_ms},
            "dx_bh_kernels": {"ms": bh_ms}}
    if world > 1:
        kern["all_gather_pvalues"] = {"ms": gather_ms, "kind": gather_kind,
                                      "note": "rank 0's view; includes waiting for the slowest rank"}
        kern["per_rank_ms"] = per_rank
        if exchange is not None:
            exchange.check()
    # the roofline object describes the kernel of the step that streams HBM (K_ingest reads the whole shard once);
    # K_fit / Wald / BH work on 4 KB of statistics per gene and are bounded by fp64 issue, not by HBM
    dom = "dx_ingest_tma_kernel"       # K <= 32 groups and a 16-byte aligned shard: the TMA variant of K_ingest runs
    achieved = kern[dom]["algorithmic_bytes"] / (kern[dom]["ms"] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "share_of_step": ingest_ms / (elapsed_ms / args.steps),
                "note": "HBM-bound kernel of the step; dx_fit_grouped_kernel (%.3f ms) runs all IRLS/Newton iterations from "
                        "on-chip statistics and is fp64-issue bound" % fit_ms}
    # SURVEY 8d's per-iteration accounting for comparison: bytes a per-iteration streaming implementation would move
    mean_iter = float(niter.to(torch.float64).mean().item())
    equiv_gbs = (8.0 * nnz * mean_iter) / (elapsed_ms / args.steps * 1e-3) / 1e9

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, 64)
        v, sec, gps = run_cpu_reference(16 * cores, steps=1, warmup=0, cores=cores)
        cpu = {"value": v, "unit": "genes/s", "cores": cores, "kind": "port",
               "sample": "%d genes x %d cells, 1 pass (oracle port of batchglm's numpy schedule, Newton scale step, "
                         "one process per core)" % (gps, N_CELLS)}

    from diffxpy_b200.testing import correction as _corr
    # own kernels per step: ingest, pack, fit, wald + BH (bucket histogram, bucket scan, place, rank, chunk-min, suffix,
    # gather; the same kernels for the gathered vector of several GPUs)
    launches_per_step = 4 + (7 if world * g <= _corr.BH_KERNEL_MAX else 2) + (1 if gather_kind == "peer" else 0)
    if rank == 0:
        conv = float((flags & 1).to(torch.float64).mean().item())
        line = {
            "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_cells": n_cells, "genes_per_gpu": g, "nnz_per_gpu": nnz,
                       "density": nnz / (n_cells * g), "design_groups": k, "p_loc": p, "p_scale": ps,
                       "l2": "inputs (%.2f GB per pass) larger than L2; no explicit flush" % (8 * nnz / 1e9),
                       "frac_converged": conv, "mean_niter": mean_iter, "max_niter": int(niter.max().item()),
                       "equivalent_iter_stream_gbs": equiv_gbs, "launch": launch_kind, "gather": gather_kind},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "genes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "dx_wald_grouped_host (C ABI, pinned host CSC in, per-gene rows out)",
                    "matches_device_path": same},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "kernels": kern,
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--genes", type=int, default=N_GENES, help="genes per GPU (default: the BASELINE workload)")
    ap.add_argument("--cells", type=int, default=N_CELLS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return main_reference(args)
    return main_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
