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


def _reference_interp(means3D, scales, rotations, opacity, shs, render_indices, parent_indices, ts, skybox):
    """The `interp_python` block of the reference's render_post (gaussian_renderer/__init__.py:278-318) as the op
    chain it is there, written against plain tensors (test-side restatement: gather, two products, one sum per
    attribute, the parent's quaternion negated when the dot product is negative, skybox rows appended)."""
    c, p = render_indices.long(), parent_indices.long()
    t, ti = ts.unsqueeze(1), (1 - ts).unsqueeze(1)
    par_rot = rotations[p]
    dots = (rotations[c] * par_rot).sum(1)
    par_rot = torch.where((dots < 0).unsqueeze(1), -par_rot, par_rot)
    n = means3D.size(0)
    tail = torch.arange(n - skybox, n, device=means3D.device)
    cat = lambda a, b: torch.cat((a, b)).contiguous()  # noqa: E731
    return (cat(t * means3D[c] + ti * means3D[p], means3D[tail]), cat(t * scales[c] + ti * scales[p], scales[tail]),
            cat(t * rotations[c] + ti * par_rot, rotations[tail]), cat(t * opacity[c] + ti * opacity[p], opacity[tail]),
            cat(t.unsqueeze(2) * shs[c] + ti.unsqueeze(2) * shs[p], shs[tail]))


@pytest.mark.gpu
@pytest.mark.parametrize("skybox", [0, 37])
def test_hierarchy_interpolation_kernel(cuda_device, skybox):
    """hg_hier_interpolate / _backward (one gather + lerp kernel each way) against the reference's PyTorch op chain:
    forward bit-exact (every product and sum rounded separately, as the chain rounds them), gradients to child, parent
    and skybox rows within float summation order (a parent is shared by several children)."""
    from hidegs_b200 import gaussian_renderer as gr
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    N, E = 5003, 3100
    mk = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa: E731
    means, scales, rots, op, shs = mk(N, 3), mk(N, 3).abs(), mk(N, 4), torch.sigmoid(mk(N, 1)), mk(N, 16, 3)
    rots = rots / rots.norm(dim=1, keepdim=True)
    ri = torch.randperm(N - skybox, generator=g)[:E].sort()[0].to(torch.int32).to(dev)
    pi = torch.randint(0, (N - skybox) // 8, (E,), generator=g, dtype=torch.int32).to(dev)  # shared parents
    pi[::11] = -1  # root nodes: torch indexing wraps to the last row (with weight 1 in a real cut)
    ts = torch.rand(E, generator=g).to(dev)
    ts[torch.rand(E, generator=g).to(dev) < 0.5] = 1.0
    ts[::11] = 1.0
    leaves_a = [x.clone().requires_grad_(True) for x in (means, scales, rots, op, shs)]
    leaves_b = [x.clone().requires_grad_(True) for x in (means, scales, rots, op, shs)]
    out_a = gr.hierarchy_interpolate(*leaves_a, ri, pi, ts, skybox)
    out_b = _reference_interp(*leaves_b, ri, pi, ts, skybox)
    for a, b, name in zip(out_a, out_b, ("means3D", "scales", "rotations", "opacity", "shs")):
        assert a.shape == b.shape == (E + skybox,) + b.shape[1:]
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
    ws = [torch.randn(o.shape, generator=g).to(dev) for o in out_b]
    sum((o * w).sum() for o, w in zip(out_a, ws)).backward()
    sum((o * w).sum() for o, w in zip(out_b, ws)).backward()
    for a, b, name in zip(leaves_a, leaves_b, ("means3D", "scales", "rotations", "opacity", "shs")):
        err = float((a.grad - b.grad).abs().max())
        assert err <= 1e-5 * float(b.grad.abs().max()), (name, err)
    # empty cut and argument errors
    e = gr.hierarchy_interpolate(means, scales, rots, op, shs, ri[:0], pi[:0], ts[:0], 0)
    assert all(o.size(0) == 0 for o in e)
    with pytest.raises(RuntimeError, match="CUDA"):
        gr.hierarchy_interpolate(means.cpu(), scales, rots, op, shs, ri, pi, ts, 0)


@pytest.mark.gpu
def test_render_post_interpolated_matches_reference_chain(cuda_device):
    """render_post(interp_python=True) — now one interpolation kernel + the rasterizer with ts / kids — renders the
    same image and returns the same gradients as the reference's op chain feeding the same rasterizer call."""
    import math
    from hidegs_b200 import gaussian_renderer as gr
    from hidegs_b200.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    dev = cuda_device
    W, H, N = 320, 208, 12_000
    sc = syn.make_scene(N, seed=6, log_scale_mean=math.log(0.01 * 1920.0 / W))
    cam = syn.default_camera(W, H).to(dev)
    g = torch.Generator().manual_seed(2)
    E = 7000
    ri = torch.randperm(N, generator=g)[:E].sort()[0].to(torch.int32).to(dev)
    pi = torch.randint(0, N, (E,), generator=g, dtype=torch.int32).to(dev)
    ts = torch.rand(E, generator=g).to(dev)
    ts[::3] = 1.0
    kids = torch.randint(2, 9, (E,), generator=g, dtype=torch.int32).to(dev)

    class PC:
        active_sh_degree = max_sh_degree = 3
        skybox_points = 0

    def model():
        pc = PC()
        pc._xyz = sc["means3D"].to(dev).clone().requires_grad_(True)
        pc.get_xyz = pc._xyz
        pc.get_scaling = sc["scales"].to(dev).clone().requires_grad_(True)
        pc.get_rotation = sc["rotations"].to(dev).clone().requires_grad_(True)
        pc.get_opacity = sc["opacity"].to(dev).clone().requires_grad_(True)
        pc.get_features = sc["shs"].to(dev).clone().requires_grad_(True)
        return pc

    class Pipe:
        compute_cov3D_python = convert_SHs_python = debug = False
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    w = torch.randn(3, H, W, generator=g).to(dev)
    a = model()
    pkg = gr.render_post(cam, a, Pipe(), bg, render_indices=ri, parent_indices=pi, interpolation_weights=ts,
                         num_node_kids=kids.clone())
    (pkg["render"] * w).sum().backward()
    b = model()
    m, s, r, o, f = _reference_interp(b.get_xyz, b.get_scaling, b.get_rotation, b.get_opacity, b.get_features, ri, pi, ts, 0)
    e_i = torch.empty(0, dtype=torch.int32, device=dev)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
        scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=3,
        campos=cam.camera_center, prefiltered=False, debug=False, render_indices=e_i, parent_indices=e_i,
        interpolation_weights=ts, num_node_kids=kids, do_depth=False, render_geo=False)
    img, radii, _, _, _, _ = GaussianRasterizer(rs)(means3D=m, means2D=torch.zeros_like(m, requires_grad=True), shs=f,
                                                    opacities=o, scales=s, rotations=r)
    (img.clamp(0, 1) * w).sum().backward()
    assert torch.equal(pkg["render"], img.clamp(0, 1))
    assert torch.equal(pkg["visibility_filter"], radii > 0)
    for x, y, name in ((a._xyz, b._xyz, "xyz"), (a.get_scaling, b.get_scaling, "scaling"),
                       (a.get_rotation, b.get_rotation, "rotation"), (a.get_opacity, b.get_opacity, "opacity"),
                       (a.get_features, b.get_features, "features")):
        err = float((x.grad - y.grad).abs().max())
        assert err <= 2e-5 * float(y.grad.abs().max()) + 1e-12, (name, err)


@pytest.mark.gpu
def test_render_coarse_and_python_sh_path(cuda_device):
    """render_coarse (reference gaussian_renderer/__init__.py:376-488): colour-only render, boolean visibility over
    all Gaussians also with a row subset; pipe.convert_SHs_python evaluates the SH colours with torch ops
    (utils/sh_utils.py eval_sh) and must render what the in-rasterizer SH evaluation renders."""
    import math
    from hidegs_b200 import gaussian_renderer as gr
    dev = cuda_device
    W, H, N = 320, 208, 9_000
    sc = syn.make_scene(N, seed=8, log_scale_mean=math.log(0.01 * 1920.0 / W))
    cam = syn.default_camera(W, H).to(dev)

    class PC:
        active_sh_degree = max_sh_degree = 3
        skybox_points = 0
    pc = PC()
    pc._xyz = pc.get_xyz = sc["means3D"].to(dev)
    pc.get_scaling, pc.get_rotation = sc["scales"].to(dev), sc["rotations"].to(dev)
    pc.get_opacity, pc.get_features = sc["opacity"].to(dev), sc["shs"].to(dev).requires_grad_(True)

    class Pipe:
        compute_cov3D_python = convert_SHs_python = debug = False

    class PipePy(Pipe):
        convert_SHs_python = True
    bg = torch.tensor([0.3, 0.1, 0.2], device=dev)
    full = gr.render(cam, pc, Pipe(), bg)
    coarse = gr.render_coarse(cam, pc, Pipe(), bg)
    assert torch.equal(coarse["render"].clamp(0, 1), full["render"])  # render() clamps (:176), render_coarse does not
    assert coarse["visibility_filter"].dtype == torch.bool and coarse["visibility_filter"].shape == (N,)
    assert torch.equal(coarse["visibility_filter"].nonzero().flatten(), full["visibility_filter"])
    assert torch.equal(coarse["radii"], full["radii"])
    idx = torch.arange(0, N, 3, device=dev)
    sub = gr.render_coarse(cam, pc, Pipe(), bg, indices=idx)
    assert sub["visibility_filter"].shape == (N,) and not sub["visibility_filter"][1::3].any()
    assert int(sub["visibility_filter"].sum()) == sub["radii"].numel() > 0
    py = gr.render_coarse(cam, pc, PipePy(), bg)
    assert float((py["render"] - coarse["render"]).abs().max()) <= 1e-5
    py["render"].sum().backward()
    g_py = pc.get_features.grad.clone()
    pc.get_features.grad = None
    gr.render_coarse(cam, pc, Pipe(), bg)["render"].sum().backward()
    ref = pc.get_features.grad
    assert float((g_py - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
