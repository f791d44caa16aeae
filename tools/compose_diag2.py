#!/usr/bin/env python
"""Diagnostic 2: where the normal-term path of the composed step departs from the oracle composition."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import test_trainer_gpu as T  # noqa: E402
import raster_utils as ru  # noqa: E402
import bench  # noqa: E402
from hidegs_b200 import gaussian_renderer as gr, trainer as tr, loss_utils as lu  # noqa: E402
from hidegs_b200.diff_gaussian_rasterization import _C as OUR  # noqa: E402
from oracle import geometry_oracle as go  # noqa: E402
import oracle.loss_oracle as lo  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    dev = torch.device("cuda", 0)
    sc, cams, gts, _ = T._setup(dev)
    cam, gt = cams[0], gts[0]
    H, W = gt.shape[-2:]
    bg = torch.zeros(3, device=dev)
    params = tr.GaussianParams.from_scene(sc, dev)
    raw = {k: v.detach().clone() for k, v in params.leaves.items()}
    xyz = raw["xyz"]
    opacity, scaling = torch.sigmoid(raw["opacity"]), torch.exp(raw["scaling"])
    rotation = torch.nn.functional.normalize(raw["rotation"])
    e_i = torch.empty(0, dtype=torch.int32, device=dev)
    e_f = torch.empty(0, dtype=torch.float32, device=dev)
    am_o = go.input_all_map(xyz, scaling, rotation, cam.world_view_transform, cam.camera_center)
    am_l = gr.geometry_all_map(xyz, scaling, rotation, cam.world_view_transform, cam.camera_center)
    print("prologue all_map: library vs oracle max abs %.3e" % float((am_l - am_o).abs().max()))
    fa = (bg, e_i, e_i, e_f, e_i, xyz, e_f, am_l, opacity, scaling, rotation, 1.0, e_f, cam.world_view_transform,
          cam.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, raw["features"], 3, cam.camera_center, False, True,
          False, True)
    f_ref = ru.ref_module().rasterize_gaussians(*fa)
    f_our = OUR.rasterize_gaussians(*fa)
    for i, n in ((1, "color"), (4, "out_all_map"), (5, "plane_depth")):
        print("%s: ours vs ref max abs %.3e  equal %s" % (n, float((f_our[i] - f_ref[i]).abs().max()), torch.equal(f_our[i], f_ref[i])))
    # upstream gradients of the normal term on each side's own maps
    iw_o = (1.0 - lo.get_img_grad_weight(gt.cpu())).clamp(0, 1) ** 2
    iw_l = (1.0 - lu.get_img_grad_weight(gt)).clamp(0, 1) ** 2
    print("image weight: library vs oracle max abs %.3e" % float((iw_l.cpu() - iw_o).abs().max()))
    K = go.intrinsic_matrix(W / (2 * math.tan(cam.FoVx / 2)), H / (2 * math.tan(cam.FoVy / 2)), 0.5 * W, 0.5 * H)
    pd_h = f_ref[5].detach().cpu().requires_grad_(True)
    am_h = f_ref[4].detach().cpu().requires_grad_(True)
    go.normal_consistency_loss(pd_h, am_h, K, iw_o, 0.015).backward()
    pd_d = f_our[5].detach().clone().requires_grad_(True)
    am_d = f_our[4].detach().clone().requires_grad_(True)
    gr.normal_consistency_loss(pd_d, am_d, cam, iw_l, 0.015).backward()
    print("upstream d/dplane_depth: kernel vs oracle rel-L2 %.3e;  d/dall_map %.3e" % (rel(pd_d.grad, pd_h.grad), rel(am_d.grad, am_h.grad)))
    # raster backward of each side on the ORACLE upstream gradients
    g = dict(color=torch.zeros(3, H, W, device=dev), all_map=am_h.grad.to(dev), plane_depth=pd_h.grad.to(dev),
             invdepth=torch.zeros(1, H, W, device=dev))
    b_ref = ru.ref_module().rasterize_gaussians_backward(*bench.bwd_tuple(fa, f_ref, g))
    b_our = OUR.rasterize_gaussians_backward(*bench.bwd_tuple(fa, f_our, g))
    for n, a, b in zip(ru.GRAD_NAMES, b_our, b_ref):
        if a is not None and a.numel():
            print("raster backward %-14s ours vs ref rel-L2 %.3e  (|ref| %.3e)" % (n, rel(a, b), float(b.norm())))
    b_ref2 = ru.ref_module().rasterize_gaussians_backward(*bench.bwd_tuple(fa, f_ref, g))
    print("raster backward ref vs ref: dL_dmeans3D %.3e dL_dall_map %.3e" % (rel(b_ref2[3], b_ref[3]), rel(b_ref2[8], b_ref[8])))
    # prologue backward on the reference's dL_dall_map
    x1, r1 = xyz.clone().requires_grad_(True), rotation.clone().requires_grad_(True)
    go.input_all_map(x1, scaling, r1, cam.world_view_transform, cam.camera_center).backward(b_ref[8])
    x2, r2 = xyz.clone().requires_grad_(True), rotation.clone().requires_grad_(True)
    gr.geometry_all_map(x2, scaling, r2, cam.world_view_transform, cam.camera_center).backward(b_ref[8])
    print("prologue backward: xyz %.3e  rotation %.3e" % (rel(x2.grad, x1.grad), rel(r2.grad, r1.grad)))


if __name__ == "__main__":
    main()
