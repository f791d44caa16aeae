from hidegs_b200.gaussian_hierarchy._C import expand_to_size, get_interpolation_weights  # noqa: F401
