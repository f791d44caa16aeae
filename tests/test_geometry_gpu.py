"""GPU parity of the render() prologue / epilogue kernels, the fused single-view normal loss, the fused Adam step and
distCUDA2: CUDA path (through the C-ABI) vs the CPU oracle, the reference-generated fixtures and, where it was built,
the UNMODIFIED reference (oracle/_ref)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import geometry_utils_t as gt
import raster_utils as ru
from oracle import geometry_oracle as go

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


class Cam:
    """Camera stub exposing what render() / get_calib_matrix_nerf read."""

    def __init__(self, K, W, H):
        self.Fx, self.Fy, self.Cx, self.Cy = K
        self.image_width, self.image_height = W, H


@pytest.mark.parametrize("name", sorted(gt.GEOMETRY_CASES))
def test_all_map_prologue(cuda_device, name):
    from hidegs_b200 import gaussian_renderer as gr
    c = gt.make_geometry_inputs(**gt.GEOMETRY_CASES[name])
    ref = np.load(os.path.join(GOLD, "geometry_ref_%s.npz" % name))
    dev = cuda_device
    xyz = c["xyz"].to(dev).requires_grad_(True)
    rot = c["rotation"].to(dev).requires_grad_(True)
    am = gr.geometry_all_map(xyz, c["scaling"].to(dev), rot, c["view"].to(dev), c["campos"].to(dev))
    (am * c["g_all_map"].to(dev)).sum().backward()
    # tolerance: fp32 op-order differences of a 3x3 transform (|values| <= ~10)
    assert np.abs(am.detach().cpu().numpy() - ref["all_map"]).max() <= 2e-5
    assert np.array_equal(am.detach().cpu().numpy()[:, 3], np.ones(am.size(0), np.float32))
    assert rel(xyz.grad.cpu().numpy(), ref["all_map_grad_xyz"]) <= 1e-5
    assert rel(rot.grad.cpu().numpy(), ref["all_map_grad_rot"]) <= 1e-5


@pytest.mark.parametrize("name", sorted(gt.GEOMETRY_CASES))
def test_depth_normal_epilogue(cuda_device, name):
    from hidegs_b200 import gaussian_renderer as gr
    p = gt.GEOMETRY_CASES[name]
    c = gt.make_geometry_inputs(**p)
    ref = np.load(os.path.join(GOLD, "geometry_ref_%s.npz" % name))
    dev = cuda_device
    cam = Cam(c["K"], p["W"], p["H"])
    depth = c["depth"].to(dev).requires_grad_(True)
    n = gr.render_normal(cam, depth)
    assert np.abs(n.detach().cpu().numpy() - ref["render_normal"]).max() <= 1e-5
    dn = gr._DepthNormal.apply(depth, c["alpha"].to(dev)[None], gr.camera_intrinsics(cam))
    assert np.abs(dn.detach().cpu().numpy() - ref["depth_normal"]).max() <= 1e-5
    (dn * c["g_normal"].to(dev)).sum().backward()
    # gradients within 1e-3 relative of the reference's autograd (north_star tolerance for gradients)
    assert rel(depth.grad.cpu().numpy(), ref["depth_normal_grad"]) <= 1e-3


@pytest.mark.parametrize("name", sorted(gt.GEOMETRY_CASES))
def test_render_normal_with_offsets_vs_reference_golden(cuda_device, name):
    """render_normal(offset=...) (gaussian_renderer/__init__.py:21-33, graphics_utils.py:130-150): normals and both
    gradients against the reference's own normal_from_depth_image (tests/golden/api_extras_ref.npz)."""
    from hidegs_b200 import gaussian_renderer as gr
    p = gt.GEOMETRY_CASES[name]
    c = gt.make_geometry_inputs(**p)
    ref = np.load(os.path.join(GOLD, "api_extras_ref.npz"))
    dev = cuda_device
    depth = c["depth"].to(dev).requires_grad_(True)
    off = gt.sample_offsets(p["H"], p["W"], p["seed"]).to(dev).requires_grad_(True)
    n = gr.render_normal(Cam(c["K"], p["W"], p["H"]), depth, offset=off)
    # 3e-4: the normals are normalised cross products of DIFFERENCES of interpolated points (|p| ~ 5, |dp| ~ 0.01), so one
    # ulp of a point is ~1e-4 of a normal component; the same expression evaluated on the CPU agrees to 5e-6, the GPU's
    # fused multiply-adds in grid_sample move 18 of 14 k components by 1e-4 .. 1.5e-4
    assert np.abs(n.detach().cpu().numpy() - ref["normal_offset_%s" % name]).max() <= 3e-4
    (n * c["g_normal"].to(dev)).sum().backward()
    assert rel(depth.grad.cpu().numpy(), ref["normal_offset_%s_grad_depth" % name]) <= 1e-3
    assert rel(off.grad.cpu().numpy(), ref["normal_offset_%s_grad_offset" % name]) <= 1e-3


def test_depth_normal_full_size_vs_oracle(cuda_device):
    from hidegs_b200 import gaussian_renderer as gr
    H, W = 1080, 1920
    g = torch.Generator().manual_seed(5)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 5.0 + 0.001 * xx + 0.002 * yy + 0.2 * torch.sin(xx * 0.05) * torch.cos(yy * 0.04) + 0.01 * torch.rand(H, W, generator=g)
    alpha = torch.rand(H, W, generator=g)
    fx = W / (2 * math.tan(math.radians(60) / 2))
    K = (fx, fx, 0.5 * W, 0.5 * H)
    want = go.depth_normal(depth[None], alpha[None], go.intrinsic_matrix(*K)).numpy()
    got = gr._DepthNormal.apply(depth.to(cuda_device), alpha.to(cuda_device)[None], gr.camera_intrinsics(Cam(K, W, H)))
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-5


@pytest.mark.parametrize("with_weight", [True, False])
def test_normal_consistency_loss_fused(cuda_device, with_weight):
    from hidegs_b200 import gaussian_renderer as gr
    p = gt.GEOMETRY_CASES["small"]
    c = gt.make_geometry_inputs(**p)
    dev = cuda_device
    cam = Cam(c["K"], p["W"], p["H"])
    K = go.intrinsic_matrix(*c["K"])
    iw = c["image_weight"] if with_weight else None
    d_o = c["depth"].clone().requires_grad_(True)
    am_o = c["out_all_map"].clone().requires_grad_(True)
    want = go.normal_consistency_loss(d_o[None], am_o, K, iw, 0.015)
    want.backward()
    d = c["depth"].to(dev)[None].requires_grad_(True)
    am = c["out_all_map"].to(dev).requires_grad_(True)
    got = gr.normal_consistency_loss(d, am, cam, iw.to(dev) if with_weight else None, 0.015)
    (got * 3.0).backward()
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    assert rel(d.grad.cpu().numpy()[0] / 3.0, d_o.grad.numpy()) <= 1e-3
    ga, gao = am.grad.cpu().numpy() / 3.0, am_o.grad.numpy()
    assert rel(ga[:3], gao[:3]) <= 1e-5 and not ga[3:].any()  # alpha is detached, distance channel unused


@pytest.mark.parametrize("mode", ["dense", "index", "mask"])
def test_fused_adam_matches_reference_optimizer(cuda_device, mode):
    from hidegs_b200.optim import Adam
    c = gt.make_geometry_inputs(**gt.GEOMETRY_CASES["small"])
    ref = np.load(os.path.join(GOLD, "geometry_ref_small.npz"))
    dev = cuda_device
    prm = torch.nn.Parameter(c["adam_p"].to(dev))
    opt = Adam([{"params": [prm], "lr": 1.6e-3, "name": "p"}], lr=0.0, eps=1e-15)
    for s in range(3):
        prm.grad = c["adam_g"][s].to(dev)
        if mode == "dense":
            r = torch.empty(0, dtype=torch.long, device=dev)
        elif mode == "index":
            r = c["adam_rel"][s].to(dev)
        else:
            r = torch.zeros(prm.size(0), dtype=torch.bool, device=dev)
            r[c["adam_rel"][s].to(dev)] = True
        opt.step(r)
    st = opt.state[prm]
    assert rel(prm.detach().cpu().numpy(), ref["adam_%s_p" % mode]) <= 1e-6
    assert rel(st["exp_avg"].cpu().numpy(), ref["adam_%s_m" % mode]) <= 1e-6
    assert rel(st["exp_avg_sq"].cpu().numpy(), ref["adam_%s_v" % mode]) <= 1e-6
    if mode != "dense":
        untouched = np.setdiff1d(np.arange(prm.size(0)), np.concatenate([r.numpy() for r in c["adam_rel"]]))
        assert np.array_equal(prm.detach().cpu().numpy()[untouched], c["adam_p"].numpy()[untouched])


def test_fused_adam_large_block_matches_torch(cuda_device):
    """59 floats per Gaussian, 200k rows, 5 steps: fused kernel vs torch.optim.Adam (same arithmetic as OurAdam)."""
    from hidegs_b200.optim import Adam
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(200_000, 59, generator=g)
    a = torch.nn.Parameter(p0.clone().to(cuda_device))
    b = torch.nn.Parameter(p0.clone().to(cuda_device))
    oa = Adam([a], lr=1e-3, eps=1e-15)
    ob = torch.optim.Adam([b], lr=1e-3, eps=1e-15)
    for s in range(5):
        gr_ = (torch.randn(200_000, 59, generator=g) * 0.01).to(cuda_device)
        a.grad, b.grad = gr_.clone(), gr_.clone()
        oa.step()
        ob.step()
    assert rel(a.detach().cpu().numpy(), b.detach().cpu().numpy()) <= 1e-6


def _knn_points(n, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * torch.tensor([6.0, 4.0, 3.0])
    pts[: n // 10] = pts[: n // 10] * 0.01 + 1.0  # a dense cluster
    pts[5] = pts[6]  # an exact duplicate
    return pts


@pytest.mark.parametrize("n", [1, 3, 4, 1000, 20_000])
def test_knn_vs_bruteforce_oracle(cuda_device, n):
    from hidegs_b200.simple_knn._C import distCUDA2
    pts = _knn_points(max(n, 8), 1)[:n]
    got = distCUDA2(pts.to(cuda_device)).cpu().numpy()
    want = go.dist_knn3(pts.numpy())
    if n < 4:
        assert (got > 1e37).all() or not np.isfinite(got).all()
        return
    # the fp64 emulation of fp32 fma can differ from the hardware by one ulp in rare double-rounding cases
    assert np.allclose(got, want, rtol=3e-7, atol=0)
    assert (got == want).mean() > 0.999


def test_knn_bit_exact_vs_reference(cuda_device):
    path = os.path.join(ru.REF_DIR, "ref_simple_knn_C.so")
    if not os.path.exists(path):
        pytest.skip("reference simple-knn not built (oracle/Makefile `make ref`)")
    if ru.REF_DIR not in sys.path:
        sys.path.insert(0, ru.REF_DIR)
    import ref_simple_knn_C
    from hidegs_b200.simple_knn._C import distCUDA2
    for n, seed in ((1000, 0), (200_003, 2), (1_000_000, 3)):
        pts = _knn_points(n, seed).to(cuda_device)
        want = ref_simple_knn_C.distCUDA2(pts)
        got = distCUDA2(pts)
        torch.cuda.synchronize()
        assert torch.equal(got, want), (n, int((got != want).sum()))


class _Pipe:
    compute_cov3D_python = False
    convert_SHs_python = False
    debug = False


class _Model:
    def __init__(self, sc, dev):
        self._xyz = sc["means3D"].to(dev).requires_grad_(True)
        self._scaling = sc["scales"].log().to(dev).requires_grad_(True)
        self._rotation = (sc["rotations"] * 1.3).to(dev).requires_grad_(True)
        self._opacity = torch.logit(sc["opacity"].clamp(1e-4, 1 - 1e-4)).to(dev).requires_grad_(True)
        self._features = sc["shs"].to(dev).requires_grad_(True)
        self.active_sh_degree, self.max_sh_degree, self.skybox_points = 3, 3, 0

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_features = property(lambda s: s._features)

    def params(self):
        return [self._xyz, self._scaling, self._rotation, self._opacity, self._features]


def test_render_matches_reference_pipeline(cuda_device):
    """render() end to end (prologue kernel -> rasterizer -> epilogue kernel, with autograd) against the reference
    pipeline assembled from its own parts: oracle prologue (pinned to the reference's get_normal) -> the UNMODIFIED
    reference rasterizer (oracle/_ref) -> oracle epilogue (pinned to the reference's normal_from_depth_image)."""
    if not ru.ref_available():
        pytest.skip("reference rasterizer not built")
    import bench
    from hidegs_b200 import gaussian_renderer as gr, synthetic as syn
    dev = cuda_device
    W, H, n = 320, 208, 20_000
    sc = syn.make_scene(n, seed=4, log_scale_mean=math.log(0.01 * 1920.0 / W))
    cam = syn.default_camera(W, H).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    g = torch.Generator().manual_seed(9)
    w_img, w_n, w_dn, w_d = (torch.randn(s, generator=g).to(dev) for s in ((3, H, W), (3, H, W), (3, H, W), (1, H, W)))

    def loss_of(pkg):
        return ((pkg["render"] * w_img).sum() + (pkg["rendered_normal"] * w_n).sum() + (pkg["depth_normal"] * w_dn).sum()
                + (pkg["plane_depth"] * w_d).sum() * 0.01 + pkg["rendered_distance"].sum() * 0.01 + pkg["depth"].sum() * 0.01)

    ours = _Model(sc, dev)
    pkg = gr.render(cam, ours, _Pipe(), bg)
    loss_of(pkg).backward()

    refm = _Model(sc, dev)
    e_i = torch.empty(0, dtype=torch.int32, device=dev)
    e_f = torch.empty(0, dtype=torch.float32, device=dev)
    # oracle prologue on the GPU tensors (torch ops; the oracle module is device agnostic)
    am = go.input_all_map(refm.get_xyz, refm.get_scaling, refm.get_rotation, cam.world_view_transform, cam.camera_center)
    fa = (bg, e_i, e_i, e_f, e_i, refm.get_xyz, e_f, am, refm.get_opacity, refm.get_scaling, refm.get_rotation, 1.0, e_f,
          cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, refm.get_features, 3,
          cam.camera_center, False, True, False, True)
    color, radii, obs, out_am, pdepth, inv = bench.RefAutograd.apply(ru.ref_module(), fa, fa[5], fa[19], fa[8], fa[9], fa[10], am)
    fxv = W / (2 * math.tan(cam.FoVx / 2))
    fyv = H / (2 * math.tan(cam.FoVy / 2))
    K = go.intrinsic_matrix(fxv, fyv, 0.5 * W, 0.5 * H).to(dev)
    # oracle epilogue with device tensors
    n_ref = go.render_normal(pdepth.squeeze().cpu(), K.cpu())  # value check on CPU
    ref_pkg = {"render": color.clamp(0, 1), "rendered_normal": out_am[0:3], "plane_depth": pdepth, "rendered_distance": out_am[4:5],
               "depth": inv, "depth_normal": _depth_normal_torch(pdepth.squeeze(), out_am[3:4].detach(), K)}
    loss_of(ref_pkg).backward()

    assert torch.equal(pkg["radii"], radii[radii > 0])
    assert torch.equal(pkg["visibility_filter"], (radii > 0).nonzero().flatten().long())
    for k in ("render", "rendered_normal", "plane_depth", "rendered_distance", "depth"):
        assert (pkg[k] - ref_pkg[k]).abs().max().item() <= 1e-4, k
    assert (pkg["depth_normal"].cpu() - n_ref * out_am[3:4].detach().cpu()).abs().max().item() <= 1e-4
    for a, b, name in zip(ours.params(), refm.params(), ("xyz", "scaling", "rotation", "opacity", "features")):
        ru.assert_grads_close([a.grad.cpu()], [b.grad.cpu()], what="render " + name)


def _depth_normal_torch(depth, alpha, K):
    """oracle epilogue evaluated with device tensors (same ops as oracle.geometry_oracle, device-agnostic copies)."""
    H, W = depth.shape
    dev = depth.device
    vx = torch.arange(W, dtype=torch.float32, device=dev) / (W - 1)
    vy = torch.arange(H, dtype=torch.float32, device=dev) / (H - 1)
    vy, vx = torch.meshgrid(vy, vx, indexing="ij")
    ndc = torch.stack([vx, vy, depth], dim=-1)
    inv_scale = torch.tensor([[W - 1, H - 1]], device=dev)
    cam_xy = ndc[..., :2] * inv_scale * ndc[..., 2:3]
    xyz = torch.cat([cam_xy, ndc[..., 2:3]], dim=-1) @ torch.inverse(K.t())
    n = torch.cross(xyz[1:H - 1, 2:W] - xyz[1:H - 1, 0:W - 2], xyz[0:H - 2, 1:W - 1] - xyz[2:H, 1:W - 1], dim=-1)
    n = torch.nn.functional.normalize(n, p=2, dim=-1)
    n = torch.nn.functional.pad(n.permute(2, 0, 1), (1, 1, 1, 1), mode="constant")
    return n * alpha


def test_edge_cases(cuda_device):
    """Empty inputs, minimum image sizes and argument errors of the geometry / optimiser / knn entry points."""
    from hidegs_b200 import gaussian_renderer as gr
    from hidegs_b200.optim import Adam
    from hidegs_b200.simple_knn._C import distCUDA2
    from hidegs_b200._geometry_lib import lib as G, Intrinsics
    dev = cuda_device
    e3, e4 = torch.zeros(0, 3, device=dev), torch.zeros(0, 4, device=dev)
    am = gr.geometry_all_map(e3, e3, e4, torch.eye(4, device=dev), torch.zeros(3, device=dev))
    assert am.shape == (0, 5)
    assert distCUDA2(e3).shape == (0,)
    p = torch.nn.Parameter(torch.zeros(0, 3, device=dev))
    p.grad = torch.zeros(0, 3, device=dev)
    Adam([p], lr=1e-3).step()
    # 3x3 image: one interior pixel
    d = torch.tensor([[1.0, 1.1, 1.2], [1.0, 1.2, 1.4], [1.1, 1.3, 1.6]], device=dev)
    cam = Cam((2.0, 2.0, 1.5, 1.5), 3, 3)
    n = gr.render_normal(cam, d)
    want = go.render_normal(d.cpu(), go.intrinsic_matrix(2.0, 2.0, 1.5, 1.5))
    assert (n.cpu() - want).abs().max().item() <= 1e-6 and n[:, 0].abs().sum().item() == 0
    # too small an image / bad arguments fail loudly
    out = torch.zeros(3, 2, 2, device=dev)
    rc = G().hg_depth_normal(d.data_ptr(), None, 2, 2, Intrinsics(1, 1, 1, 1), out.data_ptr(), None)
    assert rc == 1
    with pytest.raises(RuntimeError, match="no CPU path"):
        distCUDA2(torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        gr.geometry_all_map(torch.zeros(2, 3), torch.ones(2, 3), torch.ones(2, 4), torch.eye(4), torch.zeros(3))
