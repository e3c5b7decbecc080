from diffxpy_b200.testing import wald, lrt, wald_repeated, t_test, rank_test, two_sample, pairwise, versus_rest, \
    partition, continuous_1d
