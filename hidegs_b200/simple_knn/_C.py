"""`simple_knn._C` of the reference (submodules/simple-knn/ext.cpp:15-17): distCUDA2(points) -> (N,) mean squared
distance of every point to its three nearest neighbours (spatial.cu:15-26, simple_knn.cu:186-222), through the C-ABI
entry point hg_dist2_knn3 (include/hidegs_geometry.h).  CUDA tensors only."""
import torch

from .. import _lib
from .._geometry_lib import lib as _G


def distCUDA2(points):
    if not points.is_cuda:
        raise RuntimeError("hidegs_b200 simple_knn needs a CUDA tensor (there is no CPU path)")
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    p = points.detach().float().contiguous()
    N = p.size(0)
    means = torch.full((N,), 0.0, dtype=torch.float32, device=p.device)
    if N == 0:
        return means
    with torch.cuda.device(p.device):
        ws = torch.empty(_G().hg_dist2_knn3_workspace_bytes(N), dtype=torch.uint8, device=p.device)
        rc = _G().hg_dist2_knn3(p.data_ptr(), N, means.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "distCUDA2")
    return means
