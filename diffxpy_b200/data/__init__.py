from .design import design_matrix, preview_coef_names, constraint_system_from_star, constraint_matrix_from_string, \
    constraint_matrix_from_dict, view_coef_names, bin_continuous_covariate, DesignMatrix, DesignInfo
