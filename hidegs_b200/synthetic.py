"""Seeded synthetic scenes and cameras shared by tests and bench.

Cameras are built exactly as the reference builds them (scene/cameras.py:127-130,
utils/graphics_utils.py:38-85): `world_view_transform` and `full_proj_transform`
are the TRANSPOSED (row-vector convention) matrices the rasterizer expects.
All random draws use a CPU torch.Generator so that the same seed gives the same
scene on every machine.
"""
import math

import torch


def projection_matrix(znear, zfar, fovx, fovy, primx=0.5, primy=0.5):
    """getProjectionMatrix (utils/graphics_utils.py:59-85)."""
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top = tan_y * znear
    bottom = (1 - primy) * 2 * -top
    top = primy * 2 * top
    right = tan_x * znear
    left = (1 - primx) * 2 * -right
    right = primx * 2 * right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class Camera:
    """Minimal stand-in for scene.cameras.Camera (matrices only)."""

    def __init__(self, R, T, fovx, width, height, znear=0.01, zfar=100.0):
        self.image_width, self.image_height = int(width), int(height)
        self.FoVx = float(fovx)
        self.FoVy = 2.0 * math.atan(math.tan(fovx / 2) * height / width)
        Rt = torch.zeros(4, 4, dtype=torch.float64)
        Rt[:3, :3] = torch.as_tensor(R, dtype=torch.float64).t()
        Rt[:3, 3] = torch.as_tensor(T, dtype=torch.float64)
        Rt[3, 3] = 1.0
        self.world_view_transform = Rt.float().t().contiguous()
        self.projection_matrix = projection_matrix(znear, zfar, self.FoVx, self.FoVy).t().contiguous()
        self.full_proj_transform = (self.world_view_transform @ self.projection_matrix).contiguous()
        self.camera_center = torch.linalg.inv(self.world_view_transform)[3, :3].contiguous()
        self.tanfovx = math.tan(self.FoVx * 0.5)
        self.tanfovy = math.tan(self.FoVy * 0.5)
        # scene/cameras.py:93-96
        self.Fx = self.image_width / (2 * math.tan(self.FoVx / 2))
        self.Fy = self.image_height / (2 * math.tan(self.FoVy / 2))
        self.Cx, self.Cy = 0.5 * self.image_width, 0.5 * self.image_height
        self.image_name = "synthetic"

    def to(self, device):
        for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            setattr(self, k, getattr(self, k).to(device))
        return self


def look_at_camera(eye, target, up, fovx, width, height):
    """Camera at `eye` looking at `target` (OpenCV/COLMAP convention: +z forward, +y down)."""
    eye = torch.as_tensor(eye, dtype=torch.float64)
    f = torch.as_tensor(target, dtype=torch.float64) - eye
    f = f / f.norm()
    upv = torch.as_tensor(up, dtype=torch.float64)
    r = torch.linalg.cross(f, upv)
    r = r / r.norm()
    d = torch.linalg.cross(f, r)
    R_c2w = torch.stack([r, d, f], dim=1)  # columns: camera axes in world
    T = -(R_c2w.t() @ eye)
    return Camera(R_c2w, T, fovx, width, height)


def quaternion_to_matrix(q):
    """Real-first quaternion -> rotation matrix (pytorch3d convention,
    used by the reference at scene/gaussian_model.py:150-151)."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def geometry_all_map(xyz, scales, rotations, cam):
    """Per-Gaussian prologue of render() (gaussian_renderer/__init__.py:161-169,
    scene/gaussian_model.py:153-166): [view-space normal, 1, |n . p_cam|]."""
    Rm = quaternion_to_matrix(rotations)
    idx = scales.min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    normal = Rm.gather(2, idx).squeeze(2)
    to_cam = cam.camera_center.to(xyz.device) - xyz
    neg = (normal * to_cam).sum(-1) < 0.0
    normal = torch.where(neg[:, None], -normal, normal)
    W = cam.world_view_transform.to(xyz.device)
    local_normal = normal @ W[:3, :3]
    pts = xyz @ W[:3, :3] + W[3, :3]
    dist = (local_normal * pts).sum(-1).abs()
    out = torch.zeros((xyz.shape[0], 5), dtype=torch.float32, device=xyz.device)
    out[:, :3] = local_normal
    out[:, 3] = 1.0
    out[:, 4] = dist
    return out


def make_scene(n, seed=0, extent=(3.2, 1.8, 2.0), log_scale_mean=math.log(0.01), log_scale_std=0.6,
               sh_coeffs=16):
    """Config-2 style scene: returns a dict of CPU float32 tensors."""
    g = torch.Generator().manual_seed(seed)
    ext = torch.tensor(extent)
    xyz = (torch.rand(n, 3, generator=g) * 2 - 1) * ext
    scales = torch.exp(torch.randn(n, 3, generator=g) * log_scale_std + log_scale_mean)
    q = torch.randn(n, 4, generator=g)
    rotations = q / q.norm(dim=-1, keepdim=True)
    opacity = torch.sigmoid(torch.randn(n, 1, generator=g) * 1.5)
    shs = torch.randn(n, sh_coeffs, 3, generator=g) * 0.1
    shs[:, 0, :] = torch.randn(n, 3, generator=g) * 0.5
    return dict(means3D=xyz.contiguous(), scales=scales.contiguous(), rotations=rotations.contiguous(),
                opacity=opacity.contiguous(), shs=shs.contiguous())


def default_camera(width=1920, height=1080, fovx_deg=60.0, eye=(0.0, 0.0, -5.0)):
    """Config-2 camera: at (0,0,-5), identity rotation, looking down +z."""
    R = torch.eye(3)
    T = -torch.tensor(eye, dtype=torch.float32)
    return Camera(R, T, math.radians(fovx_deg), width, height)


def upstream_grads(width, height, seed=1, do_depth=True):
    """Random upstream gradients for (color, all_map, plane_depth, invdepth)."""
    g = torch.Generator().manual_seed(seed)
    return dict(color=torch.randn(3, height, width, generator=g),
                all_map=torch.randn(5, height, width, generator=g),
                plane_depth=torch.randn(1, height, width, generator=g) * 0.01,
                invdepth=torch.randn(1 if do_depth else 0, height, width, generator=g))


def raster_settings(cam, device, sh_degree=3, bg=(0.0, 0.0, 0.0), render_geo=True, do_depth=True,
                    scale_modifier=1.0, render_indices=None, parent_indices=None,
                    interpolation_weights=None, num_node_kids=None, debug=False):
    """Build the reference's 18-field settings tuple for `cam`."""
    from .diff_gaussian_rasterization import GaussianRasterizationSettings
    e_i = torch.empty(0, dtype=torch.int32, device=device)
    e_f = torch.empty(0, dtype=torch.float32, device=device)
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.tensor(bg, dtype=torch.float32, device=device), scale_modifier=scale_modifier,
        viewmatrix=cam.world_view_transform.to(device), projmatrix=cam.full_proj_transform.to(device),
        sh_degree=sh_degree, campos=cam.camera_center.to(device), prefiltered=False, debug=debug,
        render_indices=e_i if render_indices is None else render_indices,
        parent_indices=e_i if parent_indices is None else parent_indices,
        interpolation_weights=e_f if interpolation_weights is None else interpolation_weights,
        num_node_kids=e_i if num_node_kids is None else num_node_kids,
        do_depth=do_depth, render_geo=render_geo)


# ---------------------------------------------------------------- UAV-scale scenes
def make_uav_scene(n, seed=0, extent=(200.0, 112.0), height=30.0, log_scale_mean=math.log(0.08), log_scale_std=0.7,
                   sh_coeffs=16):
    """Ground slab x in [-200,200] m, y in [-112,112] m, z in [0,30] m (config 3 / 5 recipe)."""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(n, 3, generator=g)
    xyz[:, 0] = (xyz[:, 0] * 2 - 1) * extent[0]
    xyz[:, 1] = (xyz[:, 1] * 2 - 1) * extent[1]
    xyz[:, 2] = xyz[:, 2] * height
    scales = torch.exp(torch.randn(n, 3, generator=g) * log_scale_std + log_scale_mean)
    q = torch.randn(n, 4, generator=g)
    rotations = q / q.norm(dim=-1, keepdim=True)
    opacity = torch.sigmoid(torch.randn(n, 1, generator=g) * 1.5)
    shs = torch.randn(n, sh_coeffs, 3, generator=g) * 0.1
    shs[:, 0, :] = torch.randn(n, 3, generator=g) * 0.5
    return dict(means3D=xyz.contiguous(), scales=scales.contiguous(), rotations=rotations.contiguous(),
                opacity=opacity.contiguous(), shs=shs.contiguous())


def uav_camera(i, j, grid=8, width=1920, height=1080, altitude=120.0, fovx_deg=70.0, extent=(200.0, 112.0)):
    """Camera (i, j) of the grid x grid survey over the slab: nadir on even i + j, 20 degrees oblique otherwise."""
    cx = (-1 + (2 * i + 1) / grid) * extent[0] * 0.7
    cy = (-1 + (2 * j + 1) / grid) * extent[1] * 0.7
    eye = (cx, cy, altitude)
    tilt = 0.0 if (i + j) % 2 == 0 else math.tan(math.radians(20.0)) * altitude
    target = (cx + tilt, cy, 0.0)
    return look_at_camera(eye, target, (0.0, 1.0, 0.0), math.radians(fovx_deg), width, height)


def make_hierarchy(n_leaves, seed=0, fanout=4, extent=(200.0, 112.0, 30.0)):
    """Synthetic Gaussian hierarchy in the reference's run-time format (submodules/gaussianhierarchy/types.h:18-56,
    hierarchy_loader.cpp): nodes [N,7] int32 = (depth, parent, start, count_leafs, count_merged, start_children,
    count_children), boxes [N,2,4] float32 = (min xyz, extent | max xyz, 0).  Leaves (depth 0) hold one Gaussian each;
    every inner node holds one merged Gaussian; children of a node are contiguous; Gaussians are numbered in node order."""
    import numpy as np
    rng = np.random.default_rng(seed)
    pts = rng.random((n_leaves, 3), dtype=np.float32) * np.array(extent, np.float32)
    # group leaves along a space-filling-ish order, then build levels bottom-up
    order = np.lexsort((pts[:, 2], pts[:, 1] // 8, pts[:, 0] // 8))
    pts = pts[order]
    half = (rng.random((n_leaves, 3), dtype=np.float32) * 0.2 + 0.02).astype(np.float32)
    levels = [dict(lo=pts - half, hi=pts + half, kids=None)]
    while levels[-1]["lo"].shape[0] > 1:
        lo, hi = levels[-1]["lo"], levels[-1]["hi"]
        n = lo.shape[0]
        groups = [(i, min(i + fanout, n)) for i in range(0, n, fanout)]
        nlo = np.stack([lo[a:b].min(0) for a, b in groups])
        nhi = np.stack([hi[a:b].max(0) for a, b in groups])
        levels.append(dict(lo=nlo, hi=nhi, kids=groups))
    # number nodes top-down (root first) so that children are contiguous
    sizes = [lv["lo"].shape[0] for lv in levels]
    base = np.cumsum([0] + sizes[::-1])[:-1][::-1]  # level l starts at base[l]
    N = sum(sizes)
    nodes = np.zeros((N, 7), np.int32)
    boxes = np.zeros((N, 2, 4), np.float32)
    for l, lv in enumerate(levels):
        ids = base[l] + np.arange(sizes[l])
        boxes[ids, 0, :3], boxes[ids, 1, :3] = lv["lo"], lv["hi"]
        boxes[ids, 0, 3] = (lv["hi"] - lv["lo"]).max(1)
        nodes[ids, 0] = l
        nodes[ids, 1] = -1
        nodes[ids, 3] = 1 if l == 0 else 0
        nodes[ids, 4] = 0 if l == 0 else 1
        nodes[ids, 5] = -1
        if lv["kids"] is not None:
            for j, (a, b) in enumerate(lv["kids"]):
                nodes[ids[j], 5] = base[l - 1] + a
                nodes[ids[j], 6] = b - a
                nodes[base[l - 1] + a: base[l - 1] + b, 1] = ids[j]
    nodes[:, 2] = np.arange(N)  # one Gaussian per node, numbered in node order
    return torch.from_numpy(nodes), torch.from_numpy(boxes)
