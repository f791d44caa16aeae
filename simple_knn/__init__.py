"""Top-level alias of the reference's `simple_knn` package (scene/gaussian_model.py:21 `from simple_knn._C import
distCUDA2`)."""
from hidegs_b200.simple_knn import _C  # noqa: F401
