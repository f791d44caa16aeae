"""Top-level alias of the run-time part of the reference's `gaussian_hierarchy` package."""
from hidegs_b200.gaussian_hierarchy import _C  # noqa: F401
