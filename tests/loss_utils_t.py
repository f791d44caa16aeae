"""Seeded inputs for the loss-path tests (shared by the golden generator, the oracle tests and the GPU tests)."""
import torch
import torch.nn.functional as F

# near: unsaturated level clamps (full FFT backward); far: level 0/1 clamps saturate
LOSS_CASES = {
    "near_small": dict(H=54, W=96, noise=0.002, seed=0),
    "far_small": dict(H=54, W=96, noise=0.05, seed=1),
    "near_odd": dict(H=45, W=75, noise=0.004, seed=2),
}


class GaussiansShim:
    def __init__(self, scaling):
        self.get_scaling = scaling


def make_loss_inputs(H, W, noise, seed, n_gauss=4000, patches=256):
    g = torch.Generator().manual_seed(seed)
    gt = F.avg_pool2d(torch.rand(1, 3, H, W, generator=g), 5, stride=1, padding=2)[0].clamp(0, 1)
    render = (gt + noise * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    scaling = torch.rand(n_gauss, 3, generator=g) * 0.05
    visibility = torch.arange(0, n_gauss, 2)
    pr = torch.rand(patches, 49, generator=g)
    pn = (pr + 0.2 * torch.randn(patches, 49, generator=g)).clamp(0, 1)
    pw = torch.rand(patches, 1, generator=g)
    return dict(gt=gt.contiguous(), render=render.contiguous(), scaling=scaling, visibility=visibility, patch_ref=pr,
                patch_nea=pn, patch_w=pw)
