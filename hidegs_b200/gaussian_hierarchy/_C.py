"""`gaussian_hierarchy._C` operators (ext.cpp:15-20, torch/torch_interface.cpp:18-119), same argument lists:
the run-time LOD cut (expand_to_size / get_interpolation_weights, CUDA) and the hierarchy file format
(load_hierarchy / write_hierarchy / expand_to_target; include/hidegs_hierarchy.h)."""
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


# ----------------------------------------------------------------------------------------------------------------------
# The .hier file format and the static depth cut (include/hidegs_hierarchy.h)
HIER_SYMBOLS = ("hg_hier_layout_for", "hg_hier_probe", "hg_hier_read_raw", "hg_hier_load", "hg_hier_write",
                "hg_hier_decode_device", "hg_hier_encode_device", "hg_expand_to_target")


class HierLayout(ctypes.Structure):
    """struct hg_hier_layout."""
    _fields_ = [("P", ctypes.c_int64), ("N", ctypes.c_int64), ("compressed", ctypes.c_int32), ("pos", ctypes.c_int64),
                ("rot", ctypes.c_int64), ("scale", ctypes.c_int64), ("opacity", ctypes.c_int64), ("sh", ctypes.c_int64),
                ("nodes", ctypes.c_int64), ("boxes", ctypes.c_int64), ("file_bytes", ctypes.c_int64)]


def _H():
    L = _lib.lib()
    if not getattr(L, "_hg_hier_ready", False):
        vp, i32, i64, cs = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_char_p
        lp = ctypes.POINTER(HierLayout)
        proto = {
            "hg_hier_layout_for": (ctypes.c_int, [i64, i64, i32, lp]),
            "hg_hier_probe": (ctypes.c_int, [cs, lp]),
            "hg_hier_read_raw": (ctypes.c_int, [cs, vp, i64]),
            "hg_hier_load": (ctypes.c_int, [cs, vp, vp, vp, vp, vp, vp, vp]),
            "hg_hier_write": (ctypes.c_int, [cs, i64, i64, vp, vp, vp, vp, vp, vp, vp, i32]),
            "hg_hier_decode_device": (ctypes.c_int, [vp, lp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "hg_hier_encode_device": (ctypes.c_int, [vp, lp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "hg_expand_to_target": (i64, [vp, i64, i32, vp, i64]),
        }
        for name, (res, args) in proto.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        L._hg_hier_ready = True
    return L


def _hier_outputs(lay, device):
    P, N = int(lay.P), int(lay.N)
    f = dict(dtype=torch.float32, device=device)
    return (torch.empty((P, 3), **f), torch.empty((P, 16, 3), **f), torch.empty((P, 1), **f), torch.empty((P, 3), **f),
            torch.empty((P, 4), **f), torch.empty((N, 7), dtype=torch.int32, device=device), torch.empty((N, 2, 4), **f))


def load_hierarchy(filename, device=None):
    """LoadHierarchy (torch_interface.cpp:18-50): (pos [P,3], shs [P,16,3], alpha [P,1], scales [P,3], rot [P,4],
    nodes [N,7] int32, boxes [N,2,4]).  As in the reference the tensors are CPU tensors; with `device="cuda"` the file
    image goes through pinned memory to HBM and is decoded there (half -> float, HalfNode -> Node as kernels)."""
    H = _H()
    lay = HierLayout()
    fn = str(filename).encode()
    _lib.check(H.hg_hier_probe(fn, ctypes.byref(lay)), "load_hierarchy")
    dev = torch.device("cpu" if device is None else device)
    out = _hier_outputs(lay, dev)
    if dev.type == "cpu":
        _lib.check(H.hg_hier_load(fn, *[t.data_ptr() for t in out]), "load_hierarchy")
        return out
    raw_host = torch.empty(int(lay.file_bytes), dtype=torch.uint8).pin_memory()
    _lib.check(H.hg_hier_read_raw(fn, raw_host.data_ptr(), raw_host.numel()), "load_hierarchy")
    with torch.cuda.device(dev):
        raw = raw_host.to(dev, non_blocking=True)
        _lib.check(H.hg_hier_decode_device(raw.data_ptr(), ctypes.byref(lay), *[t.data_ptr() for t in out],
                                           torch.cuda.current_stream().cuda_stream), "load_hierarchy")
        torch.cuda.current_stream().synchronize()  # the pinned image and `raw` may be released on return
    return out


def write_hierarchy(filename, pos, shs, opacities, log_scales, rotations, nodes, boxes, compressed=True):
    """WriteHierarchy (torch_interface.cpp:52-75; HierarchyWriter::write defaults to the half variant,
    hierarchy_writer.h:32).  CUDA tensors are encoded on the device and leave as one raw image; CPU tensors are
    written section by section."""
    H = _H()
    P, N = int(pos.size(0)), int(nodes.size(0))
    fn = str(filename).encode()
    ts = [pos, shs, opacities, log_scales, rotations]
    if all(t.is_cuda for t in ts + [nodes, boxes]):
        dev = pos.device
        ts = [t.detach().to(torch.float32).contiguous() for t in ts]
        nodes_c, boxes_c = nodes.to(torch.int32).contiguous(), boxes.detach().to(torch.float32).contiguous()
        lay = HierLayout()
        _lib.check(H.hg_hier_layout_for(P, N, 1 if compressed else 0, ctypes.byref(lay)), "write_hierarchy")
        with torch.cuda.device(dev):
            raw = torch.empty(int(lay.file_bytes), dtype=torch.uint8, device=dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.check(H.hg_hier_encode_device(raw.data_ptr(), ctypes.byref(lay), ts[0].data_ptr(), ts[1].data_ptr(),
                                               ts[2].data_ptr(), ts[3].data_ptr(), ts[4].data_ptr(), nodes_c.data_ptr(),
                                               boxes_c.data_ptr(), flag.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "write_hierarchy")
            if int(flag.item()):
                raise RuntimeError("Would lose information!")
            host = torch.empty(raw.numel(), dtype=torch.uint8).pin_memory()
            host.copy_(raw)
        with open(filename, "wb") as f:
            f.write(memoryview(host.numpy()))
        return
    ts = [t.detach().cpu().to(torch.float32).contiguous() for t in ts]
    nodes_c, boxes_c = nodes.cpu().to(torch.int32).contiguous(), boxes.detach().cpu().to(torch.float32).contiguous()
    _lib.check(H.hg_hier_write(fn, P, N, ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr(),
                               ts[4].data_ptr(), nodes_c.data_ptr(), boxes_c.data_ptr(), 1 if compressed else 0),
               "write_hierarchy")


def expand_to_target(nodes, target):
    """ExpandToTarget (torch_interface.cpp:77-83): the static cut at depth `target`, int32 CPU tensor of indices."""
    H = _H()
    nodes_c = nodes.cpu().to(torch.int32).contiguous()
    N = int(nodes_c.size(0))
    count = int(H.hg_expand_to_target(nodes_c.data_ptr(), N, int(target), None, 0))
    if count < 0:
        _lib.check(1, "expand_to_target")
    out = torch.empty(count, dtype=torch.int32)
    if count:
        H.hg_expand_to_target(nodes_c.data_ptr(), N, int(target), out.data_ptr(), count)
    return out
