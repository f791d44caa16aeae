"""Drop-in for the run-time part of the reference's `gaussian_hierarchy` package (submodules/gaussianhierarchy):
`gaussian_hierarchy._C.expand_to_size` and `get_interpolation_weights`.  The hierarchy builders / loaders
(load_hierarchy, write_hierarchy, expand_to_target on the CPU) are out of scope."""
from . import _C  # noqa: F401
