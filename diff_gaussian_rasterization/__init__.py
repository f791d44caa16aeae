"""Top-level alias so that `from diff_gaussian_rasterization import ...` (as in
the reference's gaussian_renderer/__init__.py:16-18) resolves to the B200
implementation when this repository is on sys.path."""
from hidegs_b200.diff_gaussian_rasterization import (  # noqa: F401
    GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _C, _RasterizeGaussians)
