"""`gaussian_hierarchy._C` run-time operators (ext.cpp:19-20, torch/torch_interface.cpp:77-119), same argument lists."""
import ctypes

import torch

from .. import _lib
from .._geometry_lib import lib as _G


def _check(t, dtype, what):
    if not t.is_cuda:
        raise RuntimeError("hidegs_b200 gaussian_hierarchy needs CUDA tensors (there is no CPU path): %s" % what)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (what, dtype, t.dtype))
    return t.contiguous()


def expand_to_size(nodes, boxes, size, viewpoint, viewdir, render_indices, parent_indices, nodes_for_render_indices):
    """ExpandToSize: fills the three index tensors in place and returns the number of rendered Gaussians."""
    nodes, boxes = _check(nodes, torch.int32, "nodes"), _check(boxes, torch.float32, "boxes")
    viewpoint = _check(viewpoint, torch.float32, "viewpoint")
    for t, n in ((render_indices, "render_indices"), (parent_indices, "parent_indices"),
                 (nodes_for_render_indices, "nodes_for_render_indices")):
        if not t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous():
            raise RuntimeError("%s must be a contiguous int32 CUDA tensor (it is written in place)" % n)
    N = nodes.size(0)
    cap = min(render_indices.numel(), parent_indices.numel(), nodes_for_render_indices.numel())
    count = ctypes.c_int32(0)
    with torch.cuda.device(nodes.device):
        ws = torch.empty(_G().hg_expand_to_size_workspace_bytes(N), dtype=torch.uint8, device=nodes.device)
        rc = _G().hg_expand_to_size(nodes.data_ptr(), boxes.data_ptr(), N, float(size), viewpoint.data_ptr(), cap,
                                    render_indices.data_ptr(), parent_indices.data_ptr(), nodes_for_render_indices.data_ptr(),
                                    None, ws.data_ptr(), ctypes.byref(count), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "expand_to_size")
    return int(count.value)


def get_interpolation_weights(indices, size, nodes, boxes, viewpoint, viewdir, ts, num_kids):
    """GetTsIndexed: `viewpoint` / `viewdir` are HOST tensors in the reference (read with data_ptr on the CPU)."""
    indices = _check(indices, torch.int32, "indices")
    nodes, boxes = _check(nodes, torch.int32, "nodes"), _check(boxes, torch.float32, "boxes")
    if not ts.is_cuda or ts.dtype != torch.float32 or not num_kids.is_cuda or num_kids.dtype != torch.int32:
        raise RuntimeError("ts / num_kids must be float32 / int32 CUDA tensors (written in place)")
    v = viewpoint.detach().float().cpu()
    n = indices.size(0)
    with torch.cuda.device(nodes.device):
        rc = _G().hg_interpolation_weights(indices.data_ptr(), n, float(size), nodes.data_ptr(), boxes.data_ptr(),
                                           float(v[0]), float(v[1]), float(v[2]), ts.data_ptr(), num_kids.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "get_interpolation_weights")
