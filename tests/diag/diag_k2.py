"""Diagnostic: two design groups, two coefficients (closed-form start exact, location model not trained)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle.nb_glm as onb
from diffxpy_b200 import _lib
import test_gpu_fast as T
_lib.require_cuda()
n, g, nb = 3000, 96, 1
rng = np.random.default_rng(100 + nb)
cond = rng.integers(0, 2, n); batch = rng.integers(0, nb, n)
y = T._simulate(rng, n, g, cond, batch, nb, 0.02, 12.0); xd = T._design(cond, batch, nb)
y[:, 5] = 0.0; y[:, 6] = 0.0; y[7, 6] = 3.0
ref = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
print("train_loc", ref["train_loc"])
a = T._run_fit(_lib, y, xd, scale_const=True, train_loc=ref["train_loc"])
b = T._run_fit(_lib, y, xd, scale_const=False, train_loc=ref["train_loc"])
scale = np.abs(b["fi"]).max(axis=(1, 2))
d = np.abs(a["fi"] - b["fi"]).max(axis=(1, 2)) / scale
dr = np.abs(a["fi"] - ref["fisher_inv"]).max(axis=(1, 2)) / scale
dg = np.abs(b["fi"] - ref["fisher_inv"]).max(axis=(1, 2)) / scale
for gi in np.argsort(-np.nan_to_num(d))[:6]:
    print("gene %d: rel dfi lane-gen %.2e lane-oracle %.2e gen-oracle %.2e | b lane %.8f gen %.8f oracle %.8f | niter %d %d %d | ll %.10f %.10f %.10f | mean %.3f" % (
        gi, d[gi], dr[gi], dg[gi], a["b"][0][gi], b["b"][0][gi], ref["b_var"][0][gi], a["ni"][gi], b["ni"][gi], ref["niter"][gi],
        a["ll"][gi], b["ll"][gi], ref["log_likelihood"][gi], y[:, gi].mean()))
