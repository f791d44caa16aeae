"""Generate tests/golden/api_extras_ref.npz from the reference's OWN code (build container only, needs /root/reference):
the branches of its API that no call site of the reference itself exercises but a user may —
  * utils/loss_utils.ssim with window sizes other than the default 11 (value and gradient),
  * utils/graphics_utils.normal_from_depth_image with per-pixel sampling offsets (render_normal(offset=...)),
    value and the gradient of a seeded linear functional w.r.t. the depth and the offsets.
Inputs are the seeded ones of tests/loss_utils_t.py / tests/geometry_utils_t.py."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geometry_utils_t as gt  # noqa: E402
import loss_utils_t as lt  # noqa: E402

REF = "/root/reference"
SSIM_WINDOWS = (3, 7, 15)


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    lu = load("ref_loss_utils", os.path.join(REF, "utils/loss_utils.py"))
    gu = load("ref_graphics_utils", os.path.join(REF, "utils/graphics_utils.py"))
    out = {}
    inp = lt.make_loss_inputs(**lt.LOSS_CASES["near_odd"])
    for w in SSIM_WINDOWS:
        r = inp["render"].clone().requires_grad_(True)
        v = lu.ssim(r, inp["gt"], window_size=w)
        g, = torch.autograd.grad(v, r)
        out["ssim_w%d" % w], out["ssim_w%d_grad" % w] = v.item(), g.numpy()
    for name, p in gt.GEOMETRY_CASES.items():
        c = gt.make_geometry_inputs(**p)
        fx, fy, cx, cy = c["K"]
        K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]]).float()
        depth = c["depth"].clone().requires_grad_(True)
        off = gt.sample_offsets(p["H"], p["W"], p["seed"]).requires_grad_(True)
        n = gu.normal_from_depth_image(depth, K, torch.eye(4), off).permute(2, 0, 1)
        gd, go_ = torch.autograd.grad((n * c["g_normal"]).sum(), (depth, off))
        out["normal_offset_%s" % name] = n.detach().numpy()
        out["normal_offset_%s_grad_depth" % name] = gd.numpy()
        out["normal_offset_%s_grad_offset" % name] = go_.numpy()
    path = os.path.join(HERE, "api_extras_ref.npz")
    np.savez_compressed(path, **out)
    print({k: (v if np.isscalar(v) else np.asarray(v).shape) for k, v in out.items()}, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
