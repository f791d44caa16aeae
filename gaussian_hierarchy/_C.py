from hidegs_b200.gaussian_hierarchy._C import (expand_to_size, expand_to_target, get_interpolation_weights,  # noqa: F401
                                                load_hierarchy, write_hierarchy)
