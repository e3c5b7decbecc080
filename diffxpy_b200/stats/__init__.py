from .stats import likelihood_ratio_test, mann_whitney_u_test, t_test_raw, t_test_moments, wald_test, \
    wald_test_chisq, two_coef_z_test, hypergeom_test
