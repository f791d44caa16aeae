"""Seeded inputs for the geometry prologue / epilogue / Adam tests (shared by the golden generator, the oracle tests
and the GPU tests)."""
import math

import torch

GEOMETRY_CASES = {
    "small": dict(n=3000, H=37, W=53, seed=0),
    "wide": dict(n=1000, H=24, W=131, seed=1),
}


def make_geometry_inputs(n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(n, 3, generator=g) * 2 - 1) * torch.tensor([3.0, 2.0, 2.0])
    scaling = torch.exp(torch.randn(n, 3, generator=g) * 0.6 + math.log(0.02))
    scaling[::17, 1] = scaling[::17, 0]  # ties between axes: first minimum wins
    q = torch.randn(n, 4, generator=g)
    rotation = q / q.norm(dim=-1, keepdim=True)
    rotation[::5] *= 1.7  # un-normalised rows: quaternion_to_matrix divides by |q|^2
    # camera: rotation about y and x, translated back
    a, b = 0.3 + 0.1 * seed, -0.2
    Ry = torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    Rx = torch.tensor([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
    R = (Ry @ Rx).float()
    T = torch.tensor([0.1, -0.2, 5.0])
    Rt = torch.eye(4)
    Rt[:3, :3] = R.t()
    Rt[:3, 3] = T
    view = Rt.t().contiguous()  # world_view_transform (row-vector convention)
    campos = torch.linalg.inv(view)[3, :3].contiguous()
    g_all_map = torch.randn(n, 5, generator=g)
    # smooth positive depth with a few discontinuities and a flat patch (degenerate-normal free)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 4.0 + 0.02 * xx + 0.03 * yy + 0.3 * torch.sin(xx * 0.3) * torch.cos(yy * 0.2) + 0.05 * torch.rand(H, W, generator=g)
    depth[H // 3:, W // 2:] += 1.5
    alpha = torch.rand(H, W, generator=g)
    g_normal = torch.randn(3, H, W, generator=g)
    fx = W / (2 * math.tan(math.radians(60) / 2))
    K = (fx, fx * 1.01, 0.5 * W, 0.5 * H)
    out_all_map = torch.randn(5, H, W, generator=g) * 0.5
    out_all_map[3] = alpha
    image_weight = torch.rand(H, W, generator=g)
    rows = 257
    adam_p = torch.randn(rows, 3, generator=g)
    adam_g = [torch.randn(rows, 3, generator=g) * (10.0 ** -s) for s in range(3)]
    adam_rel = [torch.randperm(rows, generator=g)[:rows // 3].sort()[0] for _ in range(3)]
    return dict(xyz=xyz, scaling=scaling, rotation=rotation, view=view, campos=campos, g_all_map=g_all_map, depth=depth,
                alpha=alpha, g_normal=g_normal, K=K, out_all_map=out_all_map, image_weight=image_weight, adam_p=adam_p,
                adam_g=adam_g, adam_rel=adam_rel)


def sample_offsets(H, W, seed):
    """Per-pixel sampling offsets for render_normal(offset=...): four 2-D displacements of up to +-0.8 pixel."""
    g = torch.Generator().manual_seed(seed + 77)
    return (torch.rand(H, W, 8, generator=g) - 0.5) * 1.6
