from diffxpy_b200.testing.utils import constraint_matrix_from_string, constraint_matrix_from_dict, \
    constraint_system_from_star, design_matrix, preview_coef_names, view_coef_names, bin_continuous_covariate
