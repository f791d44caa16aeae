#!/usr/bin/env python
"""Diagnostic: gradient of the single-view normal term w.r.t. plane_depth / all_map — the fused kernel and the fp32
oracle, each against the float64 oracle on the same rendered maps (conditioning of the depth-normal cross product)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import test_trainer_gpu as T  # noqa: E402
from hidegs_b200 import gaussian_renderer as gr, trainer as tr  # noqa: E402
from oracle import geometry_oracle as go  # noqa: E402
import oracle.loss_oracle as lo  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    sc, cams, gts, _ = T._setup(dev)
    cam, gt = cams[0], gts[0]
    params = tr.GaussianParams.from_scene(sc, dev)
    pkg = gr._render_impl(cam, params, tr.PipelineParams, torch.zeros(3, device=dev), _visibility_as_mask=True)
    pd, am = pkg["plane_depth"].detach(), pkg["out_all_map"].detach()
    H, W = gt.shape[-2:]
    iw = (1.0 - lo.get_img_grad_weight(gt.cpu())).clamp(0, 1) ** 2
    K = go.intrinsic_matrix(W / (2 * math.tan(cam.FoVx / 2)), H / (2 * math.tan(cam.FoVy / 2)), 0.5 * W, 0.5 * H)
    res = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        p = pd.cpu().to(dt).requires_grad_(True)
        a = am.cpu().to(dt).requires_grad_(True)
        if dt == torch.float64:  # the oracle builds float32 helper tensors: promote through default dtype
            torch.set_default_dtype(torch.float64)
        loss = go.normal_consistency_loss(p, a, K.to(dt), iw.to(dt), 0.015)
        loss.backward()
        torch.set_default_dtype(torch.float32)
        res[name] = (float(loss), p.grad.double(), a.grad.double())
    pdr, amr = pd.clone().requires_grad_(True), am.clone().requires_grad_(True)
    ours = gr.normal_consistency_loss(pdr, amr, cam, iw.to(dev), 0.015)
    ours.backward()
    res["kernel"] = (float(ours), pdr.grad.double().cpu(), amr.grad.double().cpu())
    ref = res["f64"]
    for k in ("f32", "kernel"):
        v = res[k]
        print("%-7s loss %.9g (f64 %.9g)  d/dplane_depth rel-L2 %.3e  d/dall_map rel-L2 %.3e" % (
            k, v[0], ref[0], float((v[1] - ref[1]).norm() / ref[1].norm()), float((v[2] - ref[2]).norm() / ref[2].norm())))
    print("kernel vs f32: d/dplane_depth %.3e  d/dall_map %.3e" % (
        float((res["kernel"][1] - res["f32"][1]).norm() / res["f32"][1].norm()),
        float((res["kernel"][2] - res["f32"][2]).norm() / res["f32"][2].norm())))


if __name__ == "__main__":
    main()
