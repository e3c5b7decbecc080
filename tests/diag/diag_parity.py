import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, scipy.sparse, pandas as pd
import oracle.nb_glm as onb
from diffxpy_b200.testing.tests import _fit as fit
rng = np.random.default_rng(2024)
n, g = 100_000, 10
cond = rng.integers(0, 2, n); batch = rng.integers(0, 8, n)
xd = np.concatenate([np.ones((n, 1)), cond[:, None], (batch[:, None] == np.arange(1, 8)[None])], axis=1).astype(float)
log_mu = rng.uniform(np.log(0.005), np.log(1.0), g); log_mu[0], log_mu[1] = np.log(0.002), np.log(3.0)
r = rng.uniform(0.5, 4.0, g)
beta_c = np.where(np.arange(g) % 3 == 0, rng.normal(0, 0.5, g), 0.0)
beta_b = rng.normal(0, 0.2, (8, g)); beta_b[0] = 0
mu = np.exp(log_mu[None] + cond[:, None] * beta_c[None] + beta_b[batch])
y = rng.poisson(rng.gamma(r[None], mu / r[None])).astype(np.float64)
est = fit("nb", scipy.sparse.csr_matrix(y.astype(np.float32)), pd.DataFrame(xd, columns=["c%d"%i for i in range(9)]), pd.DataFrame(np.ones((n,1)), columns=["s0"]), gene_names=["g%d"%i for i in range(g)])
ref = onb.fit_nb(y, xd, np.ones((n, 1)), method_b="newton")
print("niter gpu", est.niter, "\nniter ref", ref["niter"])
print("flags", est.error_codes, ref["error_codes"])
print("max|da| per gene", np.abs(est.a_var - ref["a_var"]).max(0))
print("db", est.b_var - ref["b_var"])
print("ll rel diff", (est.log_likelihood - ref["log_likelihood"]) / np.abs(ref["log_likelihood"]))
print("ll gpu - ll ref (abs)", est.log_likelihood - ref["log_likelihood"])
