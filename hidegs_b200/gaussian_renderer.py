"""Drop-in for the reference's gaussian_renderer package (gaussian_renderer/__init__.py): `render`,
`render_post`, `render_normal` with the same signatures and result dictionaries, built on the B200 rasterizer and
on fused kernels for the per-Gaussian prologue and the per-pixel epilogue.

  render()        gaussian_renderer/__init__.py:36-214
      prologue    :161-169 + GaussianModel.get_normal (scene/gaussian_model.py:150-166): ~10 PyTorch ops and their
                  autograd graph -> one kernel forward (`hg_geometry_all_map`), one backward
      epilogue    :200-201 render_normal(plane_depth) * rendered_alpha.detach(): ~15 full-resolution PyTorch ops
                  (utils/graphics_utils.py:17-23,108-166) -> one kernel forward, one backward
  render_post()   :217-374 (hierarchy LOD path: Python interpolation with the parent node, then the rasterizer
                  with interpolation weights / kid counts)
  render_normal() :21-33 (offset=None path)

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


def render_normal(viewpoint_cam, depth, offset=None, normal=None, scale=1):
    """render_normal (:21-33): (3, H, W) normals of the unprojected depth map."""
    if offset is not None:
        raise NotImplementedError("hidegs_b200.render_normal implements the offset=None path the reference's render() uses")
    st = max(int(scale / 2) - 1, 0)
    d = depth[st::scale, st::scale] if scale != 1 else depth
    return _DepthNormal.apply(d, None, camera_intrinsics(viewpoint_cam, scale))


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


def _colors(viewpoint_camera, pc, pipe, override_color):
    shs = colors_precomp = None
    if override_color is None:
        if pipe.convert_SHs_python:
            raise NotImplementedError("convert_SHs_python: SH evaluation runs inside the rasterizer")
        shs = pc.get_features
    else:
        colors_precomp = override_color
    return shs, colors_precomp


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

    if render_indices.size(0) != 0:
        render_inds = render_indices.long()
        if interp_python:
            num_entries = render_indices.size(0)
            interps = interpolation_weights[:num_entries].unsqueeze(1)
            interps_inv = (1 - interpolation_weights[:num_entries]).unsqueeze(1)
            parent_inds = parent_indices[:num_entries].long()
            means3D_base = (interps * means3D[render_inds] + interps_inv * means3D[parent_inds]).contiguous()
            scales_base = (interps * scales[render_inds] + interps_inv * scales[parent_inds]).contiguous()
            shs_base = (interps.unsqueeze(2) * shs[render_inds] + interps_inv.unsqueeze(2) * shs[parent_inds]).contiguous()
            parents = rotations[parent_inds]
            rots = rotations[render_inds]
            dots = torch.bmm(rots.unsqueeze(1), parents.unsqueeze(2)).flatten()
            parents[dots < 0] *= -1
            rotations_base = ((interps * rots) + interps_inv * parents).contiguous()
            opacity_base = (interps * opacity[render_inds] + interps_inv * opacity[parent_inds]).contiguous()
            if pc.skybox_points == 0:
                skybox_inds = torch.empty(0, dtype=torch.long, device=dev)
            else:
                skybox_inds = torch.arange(pc._xyz.size(0) - pc.skybox_points, pc._xyz.size(0), device=dev).long()
            means3D = torch.cat((means3D_base, means3D[skybox_inds])).contiguous()
            shs = torch.cat((shs_base, shs[skybox_inds])).contiguous()
            opacity = torch.cat((opacity_base, opacity[skybox_inds])).contiguous()
            rotations = torch.cat((rotations_base, rotations[skybox_inds])).contiguous()
            means2D = means2D[:(num_entries + pc.skybox_points)].contiguous()
            scales = torch.cat((scales_base, scales[skybox_inds])).contiguous()
            interpolation_weights = interpolation_weights.clone().detach()
            interpolation_weights[num_entries:num_entries + pc.skybox_points] = 1.0
            num_node_kids[num_entries:num_entries + pc.skybox_points] = 1
        else:
            means3D = means3D[render_inds].contiguous()
            means2D = means2D[render_inds].contiguous()
            shs = shs[render_inds].contiguous()
            opacity = opacity[render_inds].contiguous()
            scales = scales[render_inds].contiguous()
            rotations = rotations[render_inds].contiguous()
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
