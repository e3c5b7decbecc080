#!/usr/bin/env python
"""
Multi-GPU product path check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py
Every rank calls de.test.wald / de.test.lrt on the same host matrix; the Estimator shards the genes over the ranks
(no collective on the data path) and all-gathers the per-gene rows over NCCL.  The gathered result must be identical on
every rank and identical to a single-process run (computed on rank 0 with distributed="off").
"""
import os
import sys

import numpy as np
import pandas as pd
import scipy.sparse
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    import diffxpy_b200.api as de
    from diffxpy_b200.estimator import Estimator, InputDataGLM
    rng = np.random.default_rng(5)
    n, g = 6000, 333
    cond = rng.integers(0, 2, n)
    batch = rng.integers(0, 3, n)
    lam = np.exp(rng.uniform(np.log(0.05), np.log(20), g))
    y = rng.poisson(rng.gamma(1.5, lam[None, :] * np.exp(0.4 * cond)[:, None] / 1.5)).astype(np.float64)
    sd = pd.DataFrame({"condition": cond.astype(str), "batch": batch.astype(str)})
    names = ["g%d" % i for i in range(g)]
    out = {}
    for label, data in (("csc", scipy.sparse.csc_matrix(y)), ("dense", y)):
        res = de.test.wald(data=data, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                           sample_description=sd, gene_names=names)
        est = res.model_estim
        assert est._world.size == world
        assert est.a_var.shape == (4, g) and est.fisher_inv.shape == (g, 4, 4)
        out[label] = (est.a_var.copy(), est.niter.copy(), res.pval.copy(), res.qval.copy(), est._gene_range)
    np.testing.assert_array_equal(out["csc"][0], out["dense"][0])
    # identical on every rank
    t = torch.from_numpy(out["csc"][2]).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    td.all_gather(parts, t)
    for q in parts:
        assert torch.equal(torch.nan_to_num(q), torch.nan_to_num(parts[0]))
    lrt = de.test.lrt(data=scipy.sparse.csr_matrix(y), full_formula_loc="~ 1 + condition + batch",
                      reduced_formula_loc="~ 1 + batch", sample_description=sd, gene_names=names)
    assert lrt.pval.shape == (g,)
    if rank == 0:
        # single-process run of the same test
        d = de.test.wald(data=y, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition",
                         sample_description=sd, gene_names=names, distributed="off") \
            if False else None
        idata = InputDataGLM(data=y, design_loc=np.asarray(est.input_data.design_loc), design_scale=np.ones((n, 1)),
                             feature_names=names)
        single = Estimator(idata, distributed="off")
        single.initialize(); single.train_sequence("DEFAULT"); single.finalize()
        np.testing.assert_array_equal(single.a_var, out["dense"][0])
        np.testing.assert_array_equal(single.niter, out["dense"][1])
        print("dist_check OK: world %d, gene ranges cover %d genes, sharded == single-process (bit-identical), "
              "lrt median p %.3f" % (world, g, float(np.nanmedian(lrt.pval))))
    # the gather over NVLink peer memory against NCCL: several rounds (both slots, growing sequence numbers), ragged tail
    from diffxpy_b200.dist import PeerExchange
    px = PeerExchange(5000)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(100 + rank)
    for it, n_rows in enumerate((5000, 5000, 1, 4097, 0, 256, 5000)):
        src = torch.rand(n_rows, dtype=torch.float64, device="cuda", generator=gen) + it
        got = torch.full((world * n_rows,), -1.0, dtype=torch.float64, device="cuda")
        want = torch.empty(world * n_rows, dtype=torch.float64, device="cuda")
        px.allgather(src, got)
        if n_rows:
            td.all_gather_into_tensor(want, src)
        torch.cuda.synchronize()
        assert torch.equal(got, want), "peer all-gather differs from NCCL in round %d" % it
    px.check()
    px.close()
    if rank == 0:
        print("peer exchange OK: 7 rounds identical to NCCL all_gather")
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
