#!/usr/bin/env python
"""Per-phase timing of the training step (CUDA events): render forward, losses forward, backward, optimiser.
    python tools/train_probe.py [--recipe c2|uav] [--n 2000000] [--steps 5] [--views 1]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--recipe", default="c2")
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--spatial", type=int, default=0, help="store the Gaussians in Morton order")
    a = ap.parse_args()
    from hidegs_b200 import synthetic as syn, trainer as tr, gaussian_renderer as gr, loss_utils as lu
    from hidegs_b200.frequency_regularization import frequency_regularization_pyramid_scale as freg
    dev = torch.device("cuda", 0)
    W, H = bench.WIDTH, bench.HEIGHT
    if a.recipe == "uav":
        scene = syn.make_uav_scene(a.n, seed=0)
        cams = [syn.uav_camera(i, j, width=W, height=H).to(dev) for i in range(8) for j in range(8)][:a.views]
    else:
        scene = syn.make_scene(a.n, seed=0)
        cams = [bench.camera_for(r, 0).to(dev) for r in range(a.views)]
    gts = [g.to(dev) for g in bench.make_gt_images(a.views, dev)]
    params = tr.GaussianParams.from_scene(scene, dev, spatial_order=bool(a.spatial))
    t = tr.ViewShardedTrainer(params, torch.zeros(3, device=dev), start_iteration=1000)
    o = t.opt
    names = ("zero_grad", "render_fwd", "loss_fwd", "backward", "adam")
    acc = {k: 0.0 for k in names}
    ev = lambda: torch.cuda.Event(True)  # noqa: E731
    info = {}
    for s in range(a.steps + 3):
        marks = []
        e = ev(); e.record(); marks.append(("start", e))
        params.zero_grad()
        e = ev(); e.record(); marks.append(("zero_grad", e))
        for cam, gt in zip(cams, gts):
            pkg = gr.render(cam, params, t.pipe, t.bg)
            e = ev(); e.record(); marks.append(("render_fwd", e))
            image = pkg["render"]
            loss = (1.0 - o.lambda_dssim) * lu.l1_loss(image, gt) + o.lambda_dssim * (1.0 - lu.ssim(image, gt))
            loss = loss + freg(image, gt, params, None, cam, pkg["visibility_filter"], 2000)[0]
            iw = (1.0 - lu.get_img_grad_weight(gt)).clamp(0, 1) ** 2
            loss = loss + gr.normal_consistency_loss(pkg["plane_depth"], pkg["out_all_map"], cam, iw, o.single_view_weight)
            e = ev(); e.record(); marks.append(("loss_fwd", e))
            loss.backward()
            e = ev(); e.record(); marks.append(("backward", e))
            info = {"visible": int(pkg["visibility_filter"].numel())}
        t.adam.step(grad_scale=1.0 / len(cams))
        e = ev(); e.record(); marks.append(("adam", e))
        torch.cuda.synchronize()
        if s >= 3:
            for (_, a0), (k, b0) in zip(marks[:-1], marks[1:]):
                acc[k] += a0.elapsed_time(b0)
    out = {k: round(v / a.steps, 3) for k, v in acc.items()}
    out["total_ms_per_step"] = round(sum(out.values()), 3)
    out.update(info, recipe=a.recipe, gaussians=a.n, views=a.views)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
