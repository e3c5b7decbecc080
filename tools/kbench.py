#!/usr/bin/env python
"""
Kernel tuning bench: times K_ingest / K_fit of several builds of libdxb200 (tools/build_variant.sh) on the
bench.py workload (one shard generated once) and checks that every variant produces the same outputs as the
first one.  Development tool only; the numbers that count come from bench.py.

    python tools/kbench.py [--genes 20000] [--reps 10] [--what ingest,fit] name1 name2 ...   (name "product" = libdxb200.so)
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffxpy_b200 import _lib  # noqa: E402
from diffxpy_b200.device import ptr, stream_ptr  # noqa: E402
from diffxpy_b200.estimator import group_rows  # noqa: E402


def load_variant(name):
    path = _lib.LIB_PATH if name == "product" else os.path.join(ROOT, "diffxpy_b200", "variants", "libdxb200_%s.so" % name)
    lib = C.CDLL(path)
    for fn, argtypes in _lib.SIGNATURES.items():
        if not hasattr(lib, fn):
            continue
        f = getattr(lib, fn)
        f.argtypes = argtypes
        f.restype = C.c_int
    lib.dx_last_error.restype = C.c_char_p
    return lib


def call(lib, name, *args):
    rc = getattr(lib, name)(*args)
    if rc:
        raise RuntimeError("%s: %s" % (name, lib.dx_last_error().decode()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--what", default="ingest,fit")
    ap.add_argument("--flags", type=int, default=3, help="dx_group_suffstats_ex flags (1 integer counts, 2 groups fit u16)")
    ap.add_argument("--pieces", action="store_true", help="dx_group_suffstats_pieces with the shard's work list")
    ap.add_argument("--plain", action="store_true", help="use dx_group_suffstats (no gene order / integer flag)")
    ap.add_argument("--scale-batch", action="store_true", help="scale model '~1+batch' (8 dispersion coefficients, configs[2])")
    ap.add_argument("--pack", action="store_true", help="dx_pack_overflow after the ingest + dx_nb_fit_grouped_packed")
    ap.add_argument("--reserved", type=int, default=1, help="dx_fit_opts.reserved (1 = DX_OPT_SCALE_CONST: register-resident K_fit)")
    ap.add_argument("--seed", type=int, default=1000, help="shard seed (bench.py uses 1000 + rank)")
    ap.add_argument("--mu-hi", type=float, default=1.0, help="upper end of the log-uniform gene means (default: the bench workload)")
    ap.add_argument("names", nargs="+")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    bench.MU_HI = a.mu_hi
    shard, cond, batch = bench.make_shard(torch, dev, seed=a.seed, n_genes=a.genes, n_cells=a.cells)
    nnz = shard.nnz
    g = a.genes
    xd = bench.design_for(cond, batch)
    zd = np.delete(xd, 1, axis=1) if a.scale_batch else np.ones((a.cells, 1))
    p, ps = xd.shape[1], zd.shape[1]
    uniq, inv = group_rows(np.concatenate([xd, zd], axis=1))
    k = uniq.shape[0]
    xu_loc = np.ascontiguousarray(uniq[:, :p])
    xu_scale = np.ascontiguousarray(uniq[:, p:p + ps])
    n_k = np.bincount(inv, minlength=k).astype(np.float64)
    pinv = np.ascontiguousarray(np.linalg.pinv(xu_loc))
    cell_group = torch.from_numpy(inv.astype(np.uint8)).to(dev)
    d_xl, d_xs = torch.from_numpy(xu_loc).to(dev), torch.from_numpy(xu_scale).to(dev)
    d_nk, d_pinv = torch.from_numpy(n_k).to(dev), torch.from_numpy(pinv).to(dev)
    opts = _lib.DxFitOpts(max_steps=1000, update_b_freq=5, train_loc=1, train_scale=1,
                          init_a=_lib.DX_INIT_CLOSED_FORM, init_b=_lib.DX_INIT_STANDARD, newton_max_iter=100,
                          reserved=(0 if a.scale_batch else a.reserved), lltol=1e-10, newton_xtol=1e-8)
    f64 = dict(dtype=torch.float64, device=dev)
    ref = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    print("nnz %d  genes %d  groups %d  p %d" % (nnz, g, k, p))
    for name in a.names:
        # "name@g": the same library with the general K_fit kernel (reserved = 0) -- a reference for the register-resident one
        # "name:lpg=N": the lane-resident K_fit kernel with N lanes per gene (DXB200_FIT_LPG is read at every launch)
        os.environ.pop("DXB200_FIT_LPG", None)
        full_name = name
        if ":lpg=" in name:
            name, lpg = name.split(":lpg=")
            os.environ["DXB200_FIT_LPG"] = lpg
        opts.reserved = 0 if (name.endswith("@g") or a.scale_batch) else a.reserved
        lib = load_variant(name[:-2] if name.endswith("@g") else name)
        tail = torch.zeros((g, k, _lib.DX_HIST_BINS), dtype=torch.int32, device=dev)
        ovf_count = torch.zeros(g, dtype=torch.int32, device=dev)
        ovf_val = torch.zeros(nnz, dtype=torch.float32, device=dev)
        ovf_group = torch.zeros(nnz, dtype=torch.uint8, device=dev)
        a_var, b_var = torch.zeros((p, g), **f64), torch.zeros((ps, g), **f64)
        fim_inv, ll, jac = torch.zeros((g, p, p), **f64), torch.zeros(g, **f64), torch.zeros(g, **f64)
        niter, flags = torch.zeros(g, dtype=torch.int32, device=dev), torch.zeros(g, dtype=torch.int32, device=dev)

        order, is_int = shard.prepare()
        packed = torch.zeros((g, 2), dtype=torch.int32, device=dev)

        def ingest():
            if a.pieces:
                pieces, multi = shard.work_list
                call(lib, "dx_group_suffstats_pieces", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cell_group), k,
                     ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(pieces), int(pieces.shape[0]),
                     ptr(multi) if multi.shape[0] else None, int(multi.shape[0]), (a.flags if is_int else a.flags & 2), stream_ptr())
            elif a.plain or not hasattr(lib, "dx_group_suffstats_ex"):
                call(lib, "dx_group_suffstats", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cell_group), k,
                     ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), stream_ptr())
            else:
                call(lib, "dx_group_suffstats_ex", g, ptr(shard.indptr), ptr(shard.rows), ptr(shard.vals), ptr(cell_group), k,
                     ptr(tail), ptr(ovf_count), ptr(ovf_val), ptr(ovf_group), ptr(order), (a.flags if is_int else a.flags & 2), stream_ptr())
            if a.pack:
                call(lib, "dx_pack_overflow", g, k, 0, ptr(ovf_count), ptr(shard.indptr), ptr(ovf_val), ptr(ovf_group),
                     ptr(packed), stream_ptr())

        def fit():
            if a.pack:
                call(lib, "dx_nb_fit_grouped_packed", g, k, p, ps, ptr(d_xl), ptr(d_xs), None, ptr(d_nk), ptr(d_pinv), None, k,
                     ptr(tail), ptr(ovf_count), ptr(shard.indptr), ptr(ovf_val), ptr(ovf_group), ptr(packed), opts, ptr(a_var),
                     ptr(b_var), ptr(fim_inv), ptr(ll), ptr(jac), ptr(niter), ptr(flags), stream_ptr())
                return
            call(lib, "dx_nb_fit_grouped", g, k, p, ps, ptr(d_xl), ptr(d_xs), None, ptr(d_nk), ptr(d_pinv), None, k,
                 ptr(tail), ptr(ovf_count), ptr(shard.indptr), ptr(ovf_val), ptr(ovf_group), opts, ptr(a_var),
                 ptr(b_var), ptr(fim_inv), ptr(ll), ptr(jac), ptr(niter), ptr(flags), stream_ptr())

        out = [full_name]
        for what, fn in (("ingest", ingest), ("fit", fit)):
            if what not in a.what.split(",") and not (what == "ingest"):
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.reps):
                e0, e1 = ev(), ev()
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts = np.array(ts)
            extra = ""
            if what == "ingest":
                extra = " %.0f GB/s" % ((8 * nnz + 4 * g * k * 64) / (np.median(ts) * 1e-3) / 1e9)
            out.append("%s median %.4f ms min %.4f%s" % (what, np.median(ts), ts.min(), extra))
        res = {"tail": tail, "ovf_count": ovf_count, "a": a_var, "b": b_var, "fim": fim_inv, "ll": ll, "niter": niter, "flags": flags}
        if not ref:
            ref = {kk: v.clone() for kk, v in res.items()}
            out.append("mean niter %.3f max %d conv %.4f packed genes %d max ovf %d" % (
                niter.double().mean().item(), int(niter.max().item()), (flags & 1).double().mean().item(),
                int((packed[:, 0] > 0).sum().item()), int(ovf_count.max().item())))
        else:
            ok_int = all(torch.equal(ref[kk], res[kk]) for kk in ("tail", "ovf_count"))
            out.append("suffstats_equal=%s" % ok_int)
            if "fit" in a.what.split(","):
                da = (ref["a"] - res["a"]).abs().max().item()
                dll = ((ref["ll"] - res["ll"]).abs() / ref["ll"].abs().clamp(min=1)).max().item()
                dfim = ((ref["fim"] - res["fim"]).abs() / ref["fim"].abs().clamp(min=1e-300)).nan_to_num(0).max().item()
                bitwise = all(torch.equal(torch.nan_to_num(ref[kk], nan=-7.0), torch.nan_to_num(res[kk], nan=-7.0))
                              for kk in ("a", "b", "fim", "ll"))
                out.append("niter_equal=%s flags_equal=%s bitwise_equal=%s max|da|=%.2e rel dll=%.2e rel dfim=%.2e" % (
                    torch.equal(ref["niter"], res["niter"]), torch.equal(ref["flags"], res["flags"]), bitwise, da, dll, dfim))
        print(" | ".join(out), flush=True)


if __name__ == "__main__":
    main()
