"""Generate tests/golden/api_extras_ref.npz from the reference's OWN code (build container only, needs /root/reference):
the branches of its API that no call site of the reference itself exercises but a user may —
  * utils/loss_utils.ssim with window sizes other than the default 11 (value and gradient),
  * utils/graphics_utils.normal_from_depth_image with per-pixel sampling offsets (render_normal(offset=...)),
    value and the gradient of a seeded linear functional w.r.t. the depth and the offsets,
  * scripts/frequency_regularization.frequency_regularization_pyramid_scale with level counts other than 3
    (0, 1, 2 and 4: below two levels the pyramid is the image alone; above three the level-weight table runs out,
    the reference's own try block turns the frequency term into a gradient-free zero and the scale term remains).
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
FREQ_LEVELS = (0, 1, 2, 4)


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
    fr = load("ref_frequency_regularization", os.path.join(REF, "scripts/frequency_regularization.py"))
    inp = lt.make_loss_inputs(**lt.LOSS_CASES["near_small"])
    for lv in FREQ_LEVELS:
        r = inp["render"].clone().requires_grad_(True)
        scal = inp["scaling"].clone().requires_grad_(True)
        total, hf_mask, info = fr.frequency_regularization_pyramid_scale(r, inp["gt"], lt.GaussiansShim(scal), None, None,
                                                                         inp["visibility"], 2000, num_levels=lv)
        gr, gs = torch.autograd.grad(total, (r, scal), allow_unused=True)
        out["freq_lv%d_total" % lv] = total.item()
        out["freq_lv%d_grad_render" % lv] = gr.numpy() if gr is not None else np.zeros_like(r.detach().numpy())
        out["freq_lv%d_grad_scaling" % lv] = gs.numpy() if gs is not None else np.zeros_like(scal.detach().numpy())
        out["freq_lv%d_mask_pixels" % lv] = int(hf_mask.sum().item())
        out["freq_lv%d_pyramid_levels" % lv] = int(info["pyramid_levels"])
        out["freq_lv%d_freq_loss" % lv] = float(info.get("freq_loss", -1.0))
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
