"""
Synthetic negative-binomial count generator with the surface of ``batchglm.api.models.numpy.glm_nb.Simulator``
that the reference's unit tests drive (/root/reference/diffxpy/unit_test/test_single_de.py:33-44,
test_single_null.py:24-33): ``Simulator(num_observations, num_features)``, ``generate_sample_description``,
``generate_params``, ``generate_data``, attributes ``x``, ``sample_description``, ``a_var``, ``b_var``, ``nobs``.
"""
import numpy as np
import pandas as pd

from .data.design import design_matrix


class Simulator:
    def __init__(self, num_observations=1000, num_features=100):
        self.nobs = int(num_observations)
        self.nfeatures = int(num_features)
        self.sample_description = None
        self.a_var = self.b_var = self.x = None
        self.design_loc = self.design_scale = None

    def generate_sample_description(self, num_conditions=2, num_batches=4, intercept_scale=False, shuffle_assignments=False):
        n = self.nobs
        d = {}
        if num_conditions > 0:
            d["condition"] = (np.arange(n) % num_conditions).astype(str) if not shuffle_assignments \
                else np.random.randint(num_conditions, size=n).astype(str)
        if num_batches > 0:
            d["batch"] = ((np.arange(n) // max(num_conditions, 1)) % num_batches).astype(str) if not shuffle_assignments \
                else np.random.randint(num_batches, size=n).astype(str)
        self.sample_description = pd.DataFrame(d) if d else pd.DataFrame(index=np.arange(n))
        terms = " + ".join(["1"] + list(d.keys()))
        self.formula_loc = "~ " + terms
        self.formula_scale = "~ 1" if intercept_scale or not d else "~ " + terms
        if d:
            self.design_loc = design_matrix(self.sample_description, self.formula_loc)
            self.design_scale = design_matrix(self.sample_description, self.formula_scale)
        else:
            self.design_loc = np.ones((n, 1))
            self.design_scale = np.ones((n, 1))

    def generate_params(self, rand_fn_ave=None, rand_fn=None, rand_fn_loc=None, rand_fn_scale=None):
        if rand_fn_ave is None:
            rand_fn_ave = lambda shape: np.random.poisson(500, shape) + 1
        if rand_fn is None:
            rand_fn = lambda shape: np.abs(np.random.uniform(0.5, 2, shape))
        if rand_fn_loc is None:
            rand_fn_loc = rand_fn
        if rand_fn_scale is None:
            rand_fn_scale = rand_fn
        pl, ps = np.asarray(self.design_loc).shape[1], np.asarray(self.design_scale).shape[1]
        g = self.nfeatures
        self.a_var = np.concatenate([np.log(rand_fn_ave((1, g))), rand_fn_loc((pl - 1, g))], axis=0).astype(np.float64)
        self.b_var = np.concatenate([rand_fn_scale((1, g)), rand_fn_scale((ps - 1, g)) if ps > 1 else np.zeros((0, g))],
                                    axis=0).astype(np.float64)

    def generate_data(self):
        mu = np.exp(np.asarray(self.design_loc) @ self.a_var)
        r = np.exp(np.asarray(self.design_scale) @ self.b_var)
        self.x = np.random.poisson(np.random.gamma(r, mu / r)).astype(np.float64)
        return self.x

    def generate(self):
        if self.sample_description is None:
            self.generate_sample_description()
        self.generate_params()
        self.generate_data()

    @property
    def X(self):
        return self.x

    @property
    def input_data(self):
        return self.x
