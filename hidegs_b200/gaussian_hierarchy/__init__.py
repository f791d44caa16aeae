"""Drop-in for the reference's `gaussian_hierarchy` package (submodules/gaussianhierarchy) as the training /
rendering path uses it: `gaussian_hierarchy._C.expand_to_size`, `get_interpolation_weights` (run-time LOD cut, CUDA),
`load_hierarchy`, `write_hierarchy`, `expand_to_target` (the .hier file format and the static cut).  The hierarchy
BUILDERS (GaussianHierarchyCreator / Merger executables) are out of scope."""
from . import _C  # noqa: F401
