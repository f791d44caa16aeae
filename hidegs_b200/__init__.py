"""hidegs_b200 — B200-native (sm_100a) implementation of HiDeGS's differentiable
Gaussian-splatting hot path behind the reference's own Python API.

Sub-packages / modules:
  diff_gaussian_rasterization  drop-in for submodules/hierarchy-rasterizer's
                               Python package (GaussianRasterizationSettings,
                               GaussianRasterizer, rasterize_gaussians, _C)
  _lib                         ctypes binding of the C-ABI in include/*.h
  build                        in-tree nvcc build of libhidegs_b200.so
"""
__version__ = "0.1.0"
