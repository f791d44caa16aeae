"""Hierarchy LOD cut: expand_to_size / get_interpolation_weights.
CPU: properties of the numpy oracle on a synthetic hierarchy.  GPU: CUDA path vs the oracle and, bit for bit, vs the
UNMODIFIED reference kernels (oracle/_ref/ref_lod_switching.so = runtime_switching.cu behind oracle/lod_ref_shim.cu)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from hidegs_b200 import synthetic as syn
from oracle import geometry_oracle as go

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_lod_switching.so")
VIEW = np.array([100.0, 56.0, 120.0], np.float32)


def test_oracle_cut_properties():
    nodes, boxes = syn.make_hierarchy(3000, seed=2)
    n, b = nodes.numpy(), boxes.numpy()
    prev = None
    for tgt in (1e-4, 0.02, 0.05, 0.1, 0.5, 50.0):
        ri, pi, ni = go.expand_to_size(n, b, tgt, VIEW)
        # a cut never renders a node together with one of its ancestors, and covers every leaf exactly once
        rendered = set(ni.tolist())
        for node in ni[:200]:
            a = n[node, 1]
            while a != -1:
                assert a not in rendered
                a = n[a, 1]
        covered = np.zeros(n.shape[0], bool)

        def leaves_under(i):
            if n[i, 0] == 0:
                return 1
            return sum(leaves_under(c) for c in range(n[i, 5], n[i, 5] + n[i, 6]))
        if len(ni):
            assert sum(leaves_under(i) for i in ni) == 3000 or tgt >= 50.0
        assert np.array_equal(ri, n[ni, 2])  # one Gaussian per node in this synthetic hierarchy
        assert np.array_equal(pi, np.where(n[ni, 1] != -1, n[np.maximum(n[ni, 1], 0), 2], -1))
        if prev is not None:
            assert len(ri) <= prev  # coarser target -> fewer Gaussians
        prev = len(ri)
        ts, kids = go.interpolation_weights(ni, tgt, n, b, VIEW)
        assert ts.dtype == np.float32 and ((ts >= 0) & (ts <= 1)).all()
        assert np.array_equal(kids, np.where(n[ni, 1] != -1, n[np.maximum(n[ni, 1], 0), 6], 1))
    ri, _, _ = go.expand_to_size(n, b, 1e-4, VIEW)
    assert len(ri) == 3000  # finest cut = all leaves


@pytest.mark.gpu
@pytest.mark.parametrize("n_leaves,target", [(3000, 0.05), (200_000, 0.03), (200_000, 0.3), (1000, 1e-4), (1000, 1e9)])
def test_lod_cut_vs_oracle_and_reference(cuda_device, n_leaves, target):
    from hidegs_b200.gaussian_hierarchy import _C as H
    dev = cuda_device
    nodes, boxes = syn.make_hierarchy(n_leaves, seed=3)
    N = nodes.size(0)
    nd, bd = nodes.to(dev), boxes.to(dev)
    vp = torch.from_numpy(VIEW)
    ri = torch.full((N,), -7, dtype=torch.int32, device=dev)
    pi, ni = ri.clone(), ri.clone()
    cnt = H.expand_to_size(nd, bd, target, vp.to(dev), torch.zeros(3), ri, pi, ni)
    o_ri, o_pi, o_ni = go.expand_to_size(nodes.numpy(), boxes.numpy(), target, VIEW)
    assert cnt == len(o_ri)
    assert np.array_equal(ri[:cnt].cpu().numpy(), o_ri) and np.array_equal(pi[:cnt].cpu().numpy(), o_pi)
    assert np.array_equal(ni[:cnt].cpu().numpy(), o_ni)
    assert (ri[cnt:] == -7).all()  # nothing written past the cut
    ts = torch.empty(cnt, dtype=torch.float32, device=dev)
    kids = torch.empty(cnt, dtype=torch.int32, device=dev)
    H.get_interpolation_weights(ni[:cnt].contiguous(), target, nd, bd, vp, torch.zeros(3), ts, kids)
    o_ts, o_kids = go.interpolation_weights(o_ni, target, nodes.numpy(), boxes.numpy(), VIEW)
    assert np.array_equal(kids.cpu().numpy(), o_kids)
    assert np.allclose(ts.cpu().numpy(), o_ts, rtol=0, atol=2e-7)
    if not os.path.exists(REF):
        return
    L = ctypes.CDLL(REF)
    L.ref_expand_to_size.restype = ctypes.c_int
    vp_d = vp.to(dev)
    r_ri = torch.full((N,), -7, dtype=torch.int32, device=dev)
    r_pi, r_ni = r_ri.clone(), r_ri.clone()
    vptr = ctypes.c_void_p
    f = ctypes.c_float
    rc = L.ref_expand_to_size(ctypes.c_int(N), f(target), vptr(nd.data_ptr()), vptr(bd.data_ptr()), vptr(vp_d.data_ptr()),
                              f(0), f(0), f(1), vptr(r_ri.data_ptr()), vptr(r_pi.data_ptr()), vptr(r_ni.data_ptr()))
    torch.cuda.synchronize()
    assert rc == cnt
    assert torch.equal(ri[:cnt], r_ri[:cnt]) and torch.equal(pi[:cnt], r_pi[:cnt]) and torch.equal(ni[:cnt], r_ni[:cnt])
    r_ts, r_kids = torch.empty_like(ts), torch.empty_like(kids)
    if cnt:
        L.ref_get_ts_indexed(ctypes.c_int(cnt), vptr(r_ni.data_ptr()), f(target), vptr(nd.data_ptr()), vptr(bd.data_ptr()),
                             f(float(VIEW[0])), f(float(VIEW[1])), f(float(VIEW[2])), f(0), f(0), f(1), vptr(r_ts.data_ptr()),
                             vptr(r_kids.data_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(ts.view(torch.int32), r_ts.view(torch.int32)) and torch.equal(kids, r_kids)  # bit-identical


@pytest.mark.gpu
def test_lod_cut_feeds_render_post(cuda_device):
    """The cut drives render_post exactly as the reference's training script does (expand_to_size ->
    get_interpolation_weights -> render_post with interp_python=True)."""
    import math
    from hidegs_b200 import gaussian_renderer as gr
    from hidegs_b200.gaussian_hierarchy import _C as H
    dev = cuda_device
    nodes, boxes = syn.make_hierarchy(20_000, seed=4)
    N = nodes.size(0)
    g = torch.Generator().manual_seed(1)
    centre = (boxes[:, 0, :3] + boxes[:, 1, :3]) * 0.5

    class PC:
        pass
    pc = PC()
    pc._xyz = centre.to(dev)
    pc.get_xyz = pc._xyz
    pc.get_scaling = (boxes[:, 0, 3:4].clamp(0.05, 3.0) * 0.3).expand(N, 3).contiguous().to(dev)
    q = torch.randn(N, 4, generator=g)
    pc.get_rotation = (q / q.norm(dim=1, keepdim=True)).to(dev)
    pc.get_opacity = torch.sigmoid(torch.randn(N, 1, generator=g)).to(dev)
    pc.get_features = (torch.randn(N, 16, 3, generator=g) * 0.2).to(dev)
    pc.active_sh_degree = pc.max_sh_degree = 3
    pc.skybox_points = 0

    class Pipe:
        compute_cov3D_python = convert_SHs_python = debug = False
    cam = syn.look_at_camera((100.0, 56.0, 120.0), (100.0, 56.0, 0.0), (0.0, 1.0, 0.0), math.radians(70.0), 640, 360).to(dev)
    ri = torch.zeros(N, dtype=torch.int32, device=dev)
    pi, ni = torch.zeros_like(ri), torch.zeros_like(ri)
    target = 0.02
    cnt = H.expand_to_size(nodes.to(dev), boxes.to(dev), target, cam.camera_center, torch.zeros(3), ri, pi, ni)
    assert 0 < cnt < N
    ts = torch.zeros(N, dtype=torch.float32, device=dev)
    kids = torch.zeros(N, dtype=torch.int32, device=dev)
    H.get_interpolation_weights(ni[:cnt].contiguous(), target, nodes.to(dev), boxes.to(dev), cam.camera_center.cpu(),
                                torch.zeros(3), ts, kids)
    pkg = gr.render_post(cam, pc, Pipe(), torch.zeros(3, device=dev), render_indices=ri[:cnt], parent_indices=pi[:cnt],
                         interpolation_weights=ts[:cnt], num_node_kids=kids[:cnt])
    img = pkg["render"]
    assert img.shape == (3, 360, 640) and torch.isfinite(img).all() and float(img.max()) > 0
    assert pkg["visibility_filter"].numel() == cnt and int(pkg["visibility_filter"].sum()) > 0
