"""Diagnostic: general kernel <32> and lane kernel against the oracle on designs with more than 16 groups."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle.nb_glm as onb
from diffxpy_b200 import _lib
import test_gpu_fast as T

def cmp(name, res, ref):
    reg = (ref["niter"] <= 60) & (ref["error_codes"] == 1)
    ni_eq = (res["ni"] == ref["niter"])
    dll = np.abs(res["ll"] - ref["log_likelihood"]) / np.maximum(np.abs(ref["log_likelihood"]), 1)
    da = np.abs(res["a"] - ref["a_var"]).max(axis=0)
    print("%-28s reg %.2f | niter equal (reg) %.3f | max rel dll (reg) %.2e | max |da| (reg) %.2e | b of niter mismatches %s" % (
        name, reg.mean(), ni_eq[reg].mean(), dll[reg].max(), da[reg].max(),
        np.round(ref["b_var"][0][reg & ~ni_eq], 2)[:12]))

lib = _lib
lib.require_cuda()
for case in ("k24t",):
    rng = np.random.default_rng(5)
    if case == "k24t":
        n, g, nb = 3000, 96, 12
        rng = np.random.default_rng(100 + nb)
        cond = rng.integers(0, 2, n); batch = rng.integers(0, nb, n)
        y = T._simulate(rng, n, g, cond, batch, nb, 0.02, 12.0); xd = T._design(cond, batch, nb); sf = None
        y[:, 5] = 0.0; y[:, 6] = 0.0; y[7, 6] = 3.0
    elif case == "k24":
        n, g, nb = 3000, 64, 12
        cond = rng.integers(0, 2, n); batch = rng.integers(0, nb, n)
        y = T._simulate(rng, n, g, cond, batch, nb, 0.02, 12.0); xd = T._design(cond, batch, nb); sf = None
    elif case in ("k6", "k2"):
        n, g, nb = 3000, 96, (3 if case == "k6" else 1)
        rng = np.random.default_rng(100 + nb)
        cond = rng.integers(0, 2, n); batch = rng.integers(0, nb, n)
        y = T._simulate(rng, n, g, cond, batch, nb, 0.02, 12.0); xd = T._design(cond, batch, nb); sf = None
    elif case == "k16":
        n, g, nb = 3000, 64, 8
        cond = rng.integers(0, 2, n); batch = rng.integers(0, nb, n)
        y = T._simulate(rng, n, g, cond, batch, nb, 0.02, 12.0); xd = T._design(cond, batch, nb); sf = None
    else:
        n, g, nb = 6000, 64, 8
        cond = rng.integers(0, 2, n); batch = rng.integers(0, nb, n); sf = rng.choice([0.7, 1.4], size=n)
        y = T._simulate(rng, n, g, cond, batch, nb, 0.05, 20.0, sf=sf); xd = T._design(cond, batch, nb)
    lsf = None if sf is None else np.log(sf)
    ref = onb.fit_nb(y, xd, np.ones((n, 1)), size_factors=sf, method_b="newton")
    a = T._run_fit(lib, y, xd, log_sf=lsf, scale_const=True)
    b = T._run_fit(lib, y, xd, log_sf=lsf, scale_const=False)
    print(case, "K", a["k"], "p", xd.shape[1])
    cmp("  const-scale path (lane?)", a, ref)
    cmp("  general kernel", b, ref)
    os.environ["DXB200_FIT_SW"] = "32"
    c = T._run_fit(lib, y, xd, log_sf=lsf, scale_const=False)
    del os.environ["DXB200_FIT_SW"]
    cmp("  general kernel, SW=32", c, ref)
    bad = np.where((a["ni"] != ref["niter"]) & (ref["niter"] <= 60))[0]
    for gi in bad[:4]:
        print("    gene %d: niter gpu %d oracle %d | b gpu %.6f oracle %.6f | ll gpu %.10f oracle %.10f | mean %.3f" % (
            gi, a["ni"][gi], ref["niter"][gi], a["b"][0][gi], ref["b_var"][0][gi], a["ll"][gi], ref["log_likelihood"][gi], y[:, gi].mean()))
