"""Drop-in for the reference's gaussian_renderer package (gaussian_renderer/__init__.py): `render`,
`render_post`, `render_normal` with the same signatures and result dictionaries, built on the B200 rasterizer and
on fused kernels for the per-Gaussian prologue and the per-pixel epilogue.

  render()        gaussian_renderer/__init__.py:36-214
      prologue    :161-169 + GaussianModel.get_normal (scene/gaussian_model.py:150-166): ~10 PyTorch ops and their
                  autograd graph -> one kernel forward (`hg_geometry_all_map`), one backward
      epilogue    :200-201 render_normal(plane_depth) * rendered_alpha.detach(): ~15 full-resolution PyTorch ops
                  (utils/graphics_utils.py:17-23,108-166) -> one kernel forward, one backward
  render_post()   :217-374 (hierarchy LOD path: interpolation with the parent node in ONE gather kernel
                  (`hg_hier_interpolate`, reference: ~25 PyTorch ops), then the rasterizer with interpolation
                  weights / kid counts)
  render_coarse() :376-488 (colour-only render of the coarse optimisation)
  render_normal() :21-33

`pc` is duck-typed exactly as the reference uses it: get_xyz, get_opacity, get_scaling, get_rotation,
get_features, active_sh_degree, max_sh_degree, _xyz, skybox_points, get_covariance, get_exposure_from_name.
CUDA tensors only; there is no CPU fallback.
"""
import math

import torch

from . import _lib
from ._geometry_lib import Intrinsics, lib as _G
from .diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("hidegs_b200.gaussian_renderer needs CUDA tensors (there is no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError("hidegs_b200.gaussian_renderer expects float32 tensors, got %s" % t.dtype)


class _AllMap(torch.autograd.Function):
    """input_all_map of render() (:161-169): [view-space normal (3), 1, |n . p_view|] per Gaussian."""

    @staticmethod
    def forward(ctx, xyz, scaling, rotation, viewmatrix, campos):
        _need_cuda(xyz, scaling, rotation, viewmatrix, campos)
        xyz_c, sc_c, rot_c = xyz.contiguous(), scaling.contiguous(), rotation.contiguous()
        vm, cp = viewmatrix.contiguous(), campos.contiguous()
        N = xyz_c.size(0)
        out = torch.empty((N, 5), dtype=torch.float32, device=xyz.device)
        with torch.cuda.device(xyz.device):
            rc = _G().hg_geometry_all_map(xyz_c.data_ptr(), sc_c.data_ptr(), rot_c.data_ptr(), vm.data_ptr(), cp.data_ptr(),
                                          N, out.data_ptr(), _stream())
        _lib.check(rc, "geometry_all_map")
        ctx.save_for_backward(xyz_c, sc_c, rot_c, vm, cp)
        return out

    @staticmethod
    def backward(ctx, g):
        xyz, sc, rot, vm, cp = ctx.saved_tensors
        N = xyz.size(0)
        g = g.contiguous()
        d_xyz, d_rot = torch.empty_like(xyz), torch.empty_like(rot)
        with torch.cuda.device(xyz.device):
            rc = _G().hg_geometry_all_map_backward(xyz.data_ptr(), sc.data_ptr(), rot.data_ptr(), vm.data_ptr(),
                                                   cp.data_ptr(), N, g.data_ptr(), d_xyz.data_ptr(), d_rot.data_ptr(),
                                                   _stream())
        _lib.check(rc, "geometry_all_map_backward")
        return d_xyz, None, d_rot, None, None


def geometry_all_map(xyz, scaling, rotation, viewmatrix, campos):
    return _AllMap.apply(xyz, scaling, rotation, viewmatrix, campos)


def camera_intrinsics(viewpoint_cam, scale=1.0):
    """(fx, fy, cx, cy) of Camera.get_calib_matrix_nerf (scene/cameras.py:93-96,135-138), evaluated as Python
    floats and narrowed to fp32 like `torch.tensor([...]).float()` does."""
    if hasattr(viewpoint_cam, "Fx"):
        fx, fy, cx, cy = viewpoint_cam.Fx, viewpoint_cam.Fy, viewpoint_cam.Cx, viewpoint_cam.Cy
    else:
        W, H = int(viewpoint_cam.image_width), int(viewpoint_cam.image_height)
        fx = W / (2 * math.tan(viewpoint_cam.FoVx / 2))
        fy = H / (2 * math.tan(viewpoint_cam.FoVy / 2))
        cx, cy = 0.5 * W, 0.5 * H
    return Intrinsics(fx / scale, fy / scale, cx / scale, cy / scale)


class _DepthNormal(torch.autograd.Function):
    """normal_from_depth_image(depth, K, .) permuted to (3,H,W), times alpha (treated as a constant)."""

    @staticmethod
    def forward(ctx, depth, alpha, K):
        _need_cuda(depth)
        if depth.dim() != 2:
            raise RuntimeError("depth must be (H, W)")
        d = depth.contiguous()
        a = None
        if alpha is not None:
            _need_cuda(alpha)
            a = alpha.detach().reshape(d.shape).contiguous()
        H, W = d.shape
        out = torch.empty((3, H, W), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            rc = _G().hg_depth_normal(d.data_ptr(), a.data_ptr() if a is not None else None, H, W, K, out.data_ptr(),
                                      _stream())
        _lib.check(rc, "depth_normal")
        ctx.save_for_backward(d, a if a is not None else torch.empty(0, device=d.device))
        ctx.K = K
        return out

    @staticmethod
    def backward(ctx, g):
        d, a = ctx.saved_tensors
        H, W = d.shape
        g = g.contiguous()
        gd = torch.empty_like(d)
        with torch.cuda.device(d.device):
            rc = _G().hg_depth_normal_backward(d.data_ptr(), a.data_ptr() if a.numel() else None, g.data_ptr(), H, W,
                                               ctx.K, gd.data_ptr(), _stream())
        _lib.check(rc, "depth_normal_backward")
        return gd, None, None


def _normal_from_offset_samples(depth, offset, K):
    """The offset branch of depth_pcd2normal (utils/graphics_utils.py:130-150): the four neighbours of every interior
    pixel are displaced by per-pixel 2-D offsets and read from the unprojected point map by bilinear interpolation
    (grid_sample's default align_corners=False over coordinates normalised with (size - 1), as the reference does).
    No call site of the reference passes an offset (render() never does), so this branch is a composition of device
    tensor operations with autograd, not a fused kernel."""
    H, W = depth.shape
    dev = depth.device
    u = torch.arange(W, dtype=torch.float32, device=dev) / (W - 1)
    v = torch.arange(H, dtype=torch.float32, device=dev) / (H - 1)
    scale = torch.tensor([W - 1, H - 1], dtype=torch.float32, device=dev)
    uv = torch.stack(torch.broadcast_tensors(u[None, :], v[:, None]), -1)                      # (H, W, 2) in [0, 1]
    Kmat = torch.tensor([[K.fx, 0.0, K.cx], [0.0, K.fy, K.cy], [0.0, 0.0, 1.0]], dtype=torch.float32, device=dev)
    pts = torch.cat([uv * scale * depth[..., None], depth[..., None]], -1) @ torch.inverse(Kmat.t())   # camera points
    step = torch.tensor([[0.0, 1.0], [0.0, -1.0], [1.0, 0.0], [-1.0, 0.0]], device=dev)      # bottom, top, right, left
    ij = torch.stack(torch.meshgrid(torch.arange(W, device=dev), torch.arange(H, device=dev), indexing="xy"), -1)
    at = ij[1:-1, 1:-1, None, :] + step + offset.reshape(H, W, 4, 2)[1:-1, 1:-1]
    at = torch.stack([2 * at[..., 0] / (W - 1) - 1.0, 2 * at[..., 1] / (H - 1) - 1.0], -1)
    smp = torch.nn.functional.grid_sample(pts.permute(2, 0, 1)[None], at.reshape(1, -1, 1, 2), align_corners=False)
    smp = smp.permute(0, 2, 3, 1).reshape(H - 2, W - 2, 4, 3)
    n = torch.cross(smp[:, :, 2] - smp[:, :, 3], smp[:, :, 1] - smp[:, :, 0], dim=-1)
    n = torch.nn.functional.normalize(n, p=2, dim=-1)
    return torch.nn.functional.pad(n.permute(2, 0, 1), (1, 1, 1, 1))


def render_normal(viewpoint_cam, depth, offset=None, normal=None, scale=1):
    """render_normal (:21-33): (3, H, W) normals of the unprojected depth map."""
    st = max(int(scale / 2) - 1, 0)
    d = depth[st::scale, st::scale] if scale != 1 else depth
    K = camera_intrinsics(viewpoint_cam, scale)
    if offset is not None:
        _need_cuda(depth, offset)
        return _normal_from_offset_samples(d, offset[st::scale, st::scale], K)
    return _DepthNormal.apply(d, None, K)


class _NormalConsistency(torch.autograd.Function):
    """Fused value + gradient of  weight * mean(image_weight * sum_c |depth_normal_c - rendered_normal_c|)."""

    @staticmethod
    def forward(ctx, plane_depth, all_map, image_weight, K, weight):
        _need_cuda(plane_depth, all_map)
        d = plane_depth.reshape(plane_depth.shape[-2:]).contiguous()
        am = all_map.contiguous()
        H, W = d.shape
        if am.shape != (5, H, W):
            raise RuntimeError("all_map must be (5, H, W)")
        iw = None
        if image_weight is not None:
            _need_cuda(image_weight)
            iw = image_weight.detach().reshape(H, W).contiguous()
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        out = torch.empty(1, dtype=torch.float32, device=d.device)
        gd = torch.empty_like(d) if need else None
        gam = torch.empty_like(am) if need else None
        with torch.cuda.device(d.device):
            ws = torch.empty(_G().hg_normal_consistency_workspace_bytes(H, W), dtype=torch.uint8, device=d.device)
            rc = _G().hg_normal_consistency_loss(d.data_ptr(), am.data_ptr(), iw.data_ptr() if iw is not None else None, H, W,
                                                 K, float(weight), out.data_ptr(), gd.data_ptr() if need else None,
                                                 gam.data_ptr() if need else None, ws.data_ptr(), _stream())
        _lib.check(rc, "normal_consistency_loss")
        ctx.gd, ctx.gam, ctx.dshape = gd, gam, plane_depth.shape
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        gd = (g * ctx.gd).reshape(ctx.dshape) if ctx.needs_input_grad[0] else None
        gam = g * ctx.gam if ctx.needs_input_grad[1] else None
        return gd, gam, None, None, None


def normal_consistency_loss(plane_depth, out_all_map, viewpoint_cam, image_weight=None, weight=0.015):
    """Single-view geometric term of the training loss (weight: arguments/__init__.py:118):
        weight * (image_weight * (depth_normal - rendered_normal).abs().sum(0)).mean()
    with depth_normal = render_normal(plane_depth) * rendered_alpha.detach() and rendered_normal / rendered_alpha =
    out_all_map[0:3] / out_all_map[3] exactly as render() returns them; one kernel forward, one backward."""
    return _NormalConsistency.apply(plane_depth, out_all_map, image_weight, camera_intrinsics(viewpoint_cam), weight)


def _raster_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, render_geo, do_depth, render_indices,
                     parent_indices, interpolation_weights, num_siblings):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform.cuda(),
        projmatrix=viewpoint_camera.full_proj_transform.cuda(), sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center.cuda(), prefiltered=False, render_geo=render_geo, debug=pipe.debug,
        do_depth=do_depth, render_indices=render_indices, parent_indices=parent_indices,
        interpolation_weights=interpolation_weights, num_node_kids=num_siblings)


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def sh_basis(deg, dirs):
    """Real SH basis values [N, (deg+1)^2] of unit directions [N,3], in the coefficient order and with the constants
    of utils/sh_utils.py:eval_sh (:53-112), so that eval_sh(deg, sh, dirs) == (sh[..., :K] * basis[:, None, :]).sum(-1)."""
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    cols = [torch.full_like(x, _SH_C0)]
    if deg > 0:
        cols += [-_SH_C1 * y, _SH_C1 * z, -_SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [_SH_C2[0] * xy, _SH_C2[1] * yz, _SH_C2[2] * (2.0 * zz - xx - yy), _SH_C2[3] * xz, _SH_C2[4] * (xx - yy)]
        if deg > 2:
            cols += [_SH_C3[0] * y * (3 * xx - yy), _SH_C3[1] * xy * z, _SH_C3[2] * y * (4 * zz - xx - yy),
                     _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _SH_C3[4] * x * (4 * zz - xx - yy),
                     _SH_C3[5] * z * (xx - yy), _SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(cols, dim=1)


def _colors(viewpoint_camera, pc, pipe, override_color):
    """SH features or precomputed colours of a render call (gaussian_renderer/__init__.py:139-152)."""
    shs = colors_precomp = None
    if override_color is None:
        if pipe.convert_SHs_python:
            # `pipe.convert_SHs_python`: colours evaluated by differentiable torch ops and handed to the rasterizer as
            # colors_precomp (the reference's eval_sh path); the rasterizer itself still runs on the CUDA library
            feats = pc.get_features  # [N, K, 3]
            dirs = pc.get_xyz - viewpoint_camera.camera_center.to(feats.device)[None, :]
            dirs = dirs / dirs.norm(dim=1, keepdim=True)
            basis = sh_basis(pc.active_sh_degree, dirs)  # [N, k]
            rgb = (feats[:, :basis.size(1), :] * basis[:, :, None]).sum(1)
            colors_precomp = torch.clamp_min(rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color
    return shs, colors_precomp


class _HierInterp(torch.autograd.Function):
    """Parent interpolation of a hierarchy cut (the `interp_python` branch of render_post, reference :278-318) as ONE
    gather + lerp kernel (`hg_hier_interpolate`) and one scatter kernel backward, instead of ~25 PyTorch gathers, lerps
    and concatenations with their autograd nodes.  Outputs hold E + skybox rows."""

    @staticmethod
    def forward(ctx, means3D, scales, rotations, opacity, shs, render_indices, parent_indices, ts, skybox):
        _need_cuda(means3D, scales, rotations, opacity, shs, ts)
        dev = means3D.device
        m, sc, rot, op, sh = (t.contiguous() for t in (means3D, scales, rotations, opacity, shs))
        N, M = m.size(0), sh.size(1)
        E, S = int(render_indices.size(0)), int(skybox)
        ri = render_indices.to(device=dev, dtype=torch.int32).contiguous()
        pi = parent_indices[:E].to(device=dev, dtype=torch.int32).contiguous()
        t = ts[:E].detach().contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        o_m, o_sc, o_rot = torch.empty((E + S, 3), **f32), torch.empty((E + S, 3), **f32), torch.empty((E + S, 4), **f32)
        o_op, o_sh = torch.empty((E + S, 1), **f32), torch.empty((E + S, M, 3), **f32)
        with torch.cuda.device(dev):
            rc = _G().hg_hier_interpolate(m.data_ptr(), sc.data_ptr(), rot.data_ptr(), op.data_ptr(), sh.data_ptr(), N, M,
                                          ri.data_ptr(), pi.data_ptr(), t.data_ptr(), E, S, o_m.data_ptr(), o_sc.data_ptr(),
                                          o_rot.data_ptr(), o_op.data_ptr(), o_sh.data_ptr(), _stream())
        _lib.check(rc, "hier_interpolate")
        ctx.save_for_backward(rot, ri, pi, t)
        ctx.dims = (N, M, E, S)
        ctx.set_materialize_grads(False)
        return o_m, o_sc, o_rot, o_op, o_sh

    @staticmethod
    def backward(ctx, g_m, g_sc, g_rot, g_op, g_sh):
        rot, ri, pi, t = ctx.saved_tensors
        N, M, E, S = ctx.dims
        dev = rot.device
        widths = ((3,), (3,), (4,), (1,), (M, 3))
        gs = [g.contiguous() if g is not None else None for g in (g_m, g_sc, g_rot, g_op, g_sh)]
        ds = [torch.zeros((N,) + w, dtype=torch.float32, device=dev) if (g is not None and need) else None
              for g, w, need in zip(gs, widths, ctx.needs_input_grad[:5])]
        ptr = lambda x: x.data_ptr() if x is not None else None  # noqa: E731
        with torch.cuda.device(dev):
            rc = _G().hg_hier_interpolate_backward(rot.data_ptr(), N, M, ri.data_ptr(), pi.data_ptr(), t.data_ptr(), E, S,
                                                   *[ptr(g) if d is not None else None for g, d in zip(gs, ds)],
                                                   *[ptr(d) for d in ds], _stream())
        _lib.check(rc, "hier_interpolate_backward")
        return ds[0], ds[1], ds[2], ds[3], ds[4], None, None, None, None


def hierarchy_interpolate(means3D, scales, rotations, opacity, shs, render_indices, parent_indices,
                          interpolation_weights, skybox_points=0):
    """(means3D, scales, rotations, opacity, shs) of a hierarchy cut blended with their parent nodes; differentiable."""
    return _HierInterp.apply(means3D, scales, rotations, opacity, shs, render_indices, parent_indices,
                             interpolation_weights, int(skybox_points))


def _exposure(rendered_image, exposure):
    return torch.matmul(rendered_image.permute(1, 2, 0), exposure[:3, :3]).permute(2, 0, 1) + exposure[:3, 3, None, None]


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, indices=None,
           use_trained_exp=False, return_plane=True, return_depth_normal=True):
    """render() of the reference (:36-214); same arguments, same result dictionary."""
    return _render_impl(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, indices, use_trained_exp,
                        return_plane, return_depth_normal, False)


def _render_impl(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, indices=None,
                 use_trained_exp=False, return_plane=True, return_depth_normal=True, _visibility_as_mask=False):
    """Body of render().  `_visibility_as_mask` (hidegs_b200.trainer only): return `visibility_filter` as the boolean
    mask `radii > 0` and `radii` un-compacted, and leave `depth_normal` out (the fused normal term recomputes it).  The
    reference's `nonzero()` / boolean indexing each block the host until the forward blend has finished; a training step
    that only needs the SET (mask-based scale regulariser, visibility union, densification statistics) keeps launching."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    dev = pc.get_xyz.device
    e_i = torch.empty(0, dtype=torch.int32, device=dev)
    e_f = torch.empty(0, dtype=torch.float32, device=dev)
    raster_settings = _raster_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, return_plane, True, e_i, e_i,
                                       e_f, e_i)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D, means2D, opacity = pc.get_xyz, screenspace_points, pc.get_opacity
    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation
    shs, colors_precomp = _colors(viewpoint_camera, pc, pipe, override_color)

    if indices is not None:
        means3D = means3D[indices].contiguous()
        means2D = means2D[indices].contiguous()
        shs = shs[indices].contiguous()
        opacity = opacity[indices].contiguous()
        scales = scales[indices].contiguous()
        rotations = rotations[indices].contiguous()

    input_all_map = None
    if return_plane:
        if indices is not None:
            # the reference multiplies the un-indexed normals with the indexed positions here (:163-166) and fails
            raise RuntimeError("render(indices=..., return_plane=True) is not defined by the reference "
                               "(shape mismatch between pc.get_normal and means3D[indices])")
        # get_normal tests the facing with pc._xyz (no gradient through the comparison); positions are means3D
        input_all_map = geometry_all_map(means3D, pc.get_scaling, pc.get_rotation, raster_settings.viewmatrix,
                                         raster_settings.campos)

    rendered_image, radii, out_observe, out_all_map, plane_depth, depth_image = rasterizer(
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity, scales=scales,
        rotations=rotations, all_map=input_all_map, cov3D_precomp=cov3D_precomp)

    if use_trained_exp:
        rendered_image = _exposure(rendered_image, pc.get_exposure_from_name(viewpoint_camera.image_name))
    rendered_image = rendered_image.clamp(0, 1)

    subfilter = radii > 0
    if indices is not None:
        vis_filter = torch.zeros(pc._xyz.size(0), dtype=torch.bool, device=dev)
        w = vis_filter[indices]
        w[subfilter] = True
        vis_filter[indices] = w
    else:
        vis_filter = subfilter

    if not return_plane:
        return {"render": rendered_image, "plane_depth": plane_depth, "depth": depth_image, "out_observe": out_observe,
                "viewspace_points": screenspace_points, "visibility_filter": vis_filter.nonzero().flatten().long(),
                "radii": radii[subfilter]}

    rendered_normal = out_all_map[0:3]
    rendered_alpha = out_all_map[3:4]
    rendered_distance = out_all_map[4:5]
    out = {"render": rendered_image, "depth": depth_image, "viewspace_points": screenspace_points,
           "visibility_filter": vis_filter if _visibility_as_mask else vis_filter.nonzero().flatten().long(),
           "radii": radii if _visibility_as_mask else radii[subfilter], "out_observe": out_observe,
           "rendered_normal": rendered_normal, "plane_depth": plane_depth, "rendered_distance": rendered_distance}
    if _visibility_as_mask:
        pass  # the trainer's fused normal term computes the depth normal itself (normal_consistency_loss)
    elif return_depth_normal:
        out["depth_normal"] = _DepthNormal.apply(plane_depth.squeeze(), rendered_alpha,
                                                 camera_intrinsics(viewpoint_camera))
    else:
        # the reference reads `depth_normal` unconditionally (:213) and raises UnboundLocalError here
        raise UnboundLocalError("depth_normal is only defined for return_depth_normal=True (reference :200-213)")
    # kept for the fused training step (not part of the reference's dictionary contract, ignored by its callers)
    out["out_all_map"] = out_all_map
    return out


def render_post(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None,
                render_indices=torch.Tensor([]).int(), parent_indices=torch.Tensor([]).int(),
                interpolation_weights=torch.Tensor([]).float(), num_node_kids=torch.Tensor([]).int(),
                interp_python=True, use_trained_exp=False):
    """render_post() of the reference (:217-374): render a hierarchy cut."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    dev = pc.get_xyz.device
    means3D, means2D, opacity = pc.get_xyz, screenspace_points, pc.get_opacity
    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation
    shs, colors_precomp = _colors(viewpoint_camera, pc, pipe, override_color)

    E = int(render_indices.size(0))
    if E != 0:
        if interp_python:
            if shs is None or scales is None:
                raise RuntimeError("render_post(interp_python=True) blends SH features, scales and rotations with the "
                                   "parent node: it needs pc.get_features / get_scaling / get_rotation (reference :281-291)")
            S = int(pc.skybox_points)
            means3D, scales, rotations, opacity, shs = hierarchy_interpolate(
                means3D, scales, rotations, opacity, shs, render_indices, parent_indices, interpolation_weights, S)
            means2D = means2D[:E + S].contiguous()
            # the skybox tail is rendered as-is: weight 1, one kid (reference :313-315; like the reference, the caller's
            # num_node_kids is updated in place, the weights on a private copy)
            interpolation_weights = interpolation_weights.clone().detach()
            interpolation_weights[E:E + S] = 1.0
            num_node_kids[E:E + S] = 1
        else:
            sel = render_indices.long()
            means3D, means2D, opacity = means3D[sel].contiguous(), means2D[sel].contiguous(), opacity[sel].contiguous()
            shs, scales, rotations = shs[sel].contiguous(), scales[sel].contiguous(), rotations[sel].contiguous()
        render_indices = torch.empty(0, dtype=torch.int32, device=dev)
        parent_indices = torch.empty(0, dtype=torch.int32, device=dev)

    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=pipe.debug,
        render_indices=render_indices.to(dev), parent_indices=parent_indices.to(dev),
        interpolation_weights=interpolation_weights.to(dev), num_node_kids=num_node_kids.to(dev), do_depth=False,
        render_geo=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    rendered_image, radii, _, _, _, _ = rasterizer(
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity, scales=scales,
        rotations=rotations, cov3D_precomp=cov3D_precomp)
    if use_trained_exp and getattr(pc, "pretrained_exposures", None):
        try:
            rendered_image = _exposure(rendered_image, pc.pretrained_exposures[viewpoint_camera.image_name])
        except Exception:
            print(f"Exposures should be optimized in single. Missing exposure for image {viewpoint_camera.image_name}")
    rendered_image = rendered_image.clamp(0, 1)
    vis_filter = radii > 0
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": vis_filter,
            "radii": radii[vis_filter]}


def render_coarse(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, zfar=0.0, override_color=None, indices=None):
    """render_coarse() of the reference (:376-488): colour-only render for the coarse optimisation (render_geo=False,
    do_depth=False, debug=True as the reference sets it), optional row subset `indices`; returns the boolean
    visibility mask over ALL Gaussians and the radii of the visible rendered rows."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    dev = pc.get_xyz.device
    e_i = torch.empty(0, dtype=torch.int32, device=dev)
    e_f = torch.empty(0, dtype=torch.float32, device=dev)

    class _Debug:  # the reference hard-codes debug=True for this entry point (:412)
        debug = True
    rs = _raster_settings(viewpoint_camera, pc, _Debug, bg_color, scaling_modifier, False, False, e_i, e_i, e_f, e_i)
    means3D, means2D, opacity = pc.get_xyz, screenspace_points, pc.get_opacity
    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation
    shs, colors_precomp = _colors(viewpoint_camera, pc, pipe, override_color)
    if indices is not None:
        means3D, means2D, opacity = means3D[indices].contiguous(), means2D[indices].contiguous(), opacity[indices].contiguous()
        shs, scales, rotations = shs[indices].contiguous(), scales[indices].contiguous(), rotations[indices].contiguous()
    rendered_image, radii, _, _, _, _ = GaussianRasterizer(raster_settings=rs)(
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity, scales=scales,
        rotations=rotations, cov3D_precomp=cov3D_precomp)
    subfilter = radii > 0
    if indices is not None:
        vis_filter = torch.zeros(pc._xyz.size(0), dtype=torch.bool, device=dev)
        vis_filter[indices] = subfilter
    else:
        vis_filter = subfilter
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": vis_filter,
            "radii": radii[subfilter]}
