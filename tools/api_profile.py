"""Where the time of de.test.wald(host scipy CSR).summary() goes at the bench size (cProfile, cumulative)."""
import cProfile, pstats, io, os, sys, time
import numpy as np, pandas as pd, scipy.sparse, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import diffxpy_b200.api as de
dev = torch.device("cuda", 0)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
shard, cond, batch = bench.make_shard(torch, dev, seed=1000, n_genes=g, n_cells=100000)
csc = scipy.sparse.csc_matrix((shard.vals.cpu().numpy(), shard.rows.cpu().numpy(), shard.indptr.cpu().numpy()), shape=(100000, g))
csr = csc.tocsr()
del shard
torch.cuda.empty_cache()
sd = pd.DataFrame({"condition": cond.astype(str), "batch": batch.astype(str)})
names = np.array(["g%d" % i for i in range(g)])
def call(data):
    res = de.test.wald(data=data, formula_loc="~ 1 + condition + batch", factor_loc_totest="condition", sample_description=sd,
                       gene_names=names, noise_model="nb")
    s = res.summary()
    torch.cuda.synchronize()
    return s
from diffxpy_b200.device import _trace
for label, data in (("csr", csr), ("csc", csc)):
    call(data)
    os.environ["DXB200_TRACE"] = "1"
    _trace("call starts (%s)" % label); call(data); _trace("call done")
    del os.environ["DXB200_TRACE"]
    t0 = time.perf_counter(); call(data); t1 = time.perf_counter()
    print("%s: %.1f ms per call" % (label, (t1 - t0) * 1e3))
    pr = cProfile.Profile(); pr.enable(); call(data); pr.disable()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28)
    print("\n".join(l[:150] for l in st.getvalue().splitlines()[4:40]))
