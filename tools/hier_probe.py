#!/usr/bin/env python
"""configs[2] timing probe: 6M-Gaussian UAV slab, 3840x2160, synthetic hierarchy cut rendered the way render_post does
(interpolation weights + kid counts inside the rasterizer).  Times forward and backward (CUDA events, median) of this
library and, when oracle/_ref is present, of the unmodified reference on the same tensors.
   python tools/hier_probe.py [--iters 5]"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raster_utils as ru  # noqa: E402
from hidegs_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--n", type=int, default=6_000_000)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    W, H, n = 3840, 2160, a.n
    sc = syn.make_uav_scene(n, seed=0)
    cam = syn.look_at_camera((0.0, 0.0, 120.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), math.radians(70.0), W, H)
    g = torch.Generator().manual_seed(5)
    keep = (torch.rand(n, generator=g) < 0.6).nonzero().flatten()
    P = keep.numel()
    rest = torch.ones(n, dtype=torch.bool)
    rest[keep] = False
    rest = rest.nonzero().flatten()
    parents = rest[torch.randint(0, rest.numel(), (P,), generator=g)]
    ts = torch.rand(P, generator=g)
    ts[torch.rand(P, generator=g) < 0.5] = 1.0
    kids = torch.randint(2, 9, (P,), generator=g, dtype=torch.int32)
    d = {k: v.to(dev) for k, v in sc.items()}
    keep_d, par_d, ts_d, kids_d = keep.to(dev), parents.to(dev), ts.to(dev), kids.to(dev)
    e_i = torch.empty(0, dtype=torch.int32, device=dev)
    e_f = torch.empty(0, dtype=torch.float32, device=dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    camd = cam.to(dev)
    t1, t0 = ts_d[:, None], (1 - ts_d)[:, None]
    m3 = (t1 * d["means3D"][keep_d] + t0 * d["means3D"][par_d]).contiguous()
    scl = (t1 * d["scales"][keep_d] + t0 * d["scales"][par_d]).contiguous()
    shs = (t1[:, :, None] * d["shs"][keep_d] + t0[:, :, None] * d["shs"][par_d]).contiguous()
    rp, rc = d["rotations"][par_d].clone(), d["rotations"][keep_d]
    rp[(rc * rp).sum(1) < 0] *= -1
    rot = (t1 * rc + t0 * rp).contiguous()
    op = (t1 * d["opacity"][keep_d] + t0 * d["opacity"][par_d]).contiguous()
    del d
    fa = (bg, e_i, e_i, ts_d, kids_d, m3, e_f, e_f, op, scl, rot, 1.0, e_f, camd.world_view_transform,
          camd.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, shs, 3, camd.camera_center, False, False, False, False)
    gr = dict(color=torch.randn(3, H, W, generator=g).to(dev), all_map=torch.zeros(5, H, W, device=dev),
              plane_depth=torch.zeros(1, H, W, device=dev), invdepth=torch.zeros(0, H, W, device=dev))

    def time_impl(C, after_warmup=None):
        f_ms, b_ms, R = [], [], 0
        for it in range(a.iters + 2):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            torch.cuda.synchronize()
            if it == 2 and after_warmup is not None:
                after_warmup()  # (stage records of the two warm-up calls, with their one-time module loads, are dropped)
            e[0].record()
            fwd = C.rasterize_gaussians(*fa)
            e[1].record()
            bwd = C.rasterize_gaussians_backward(*ru.bwd_args(fa, fwd, gr, dev))
            e[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                f_ms.append(e[0].elapsed_time(e[1]))
                b_ms.append(e[1].elapsed_time(e[2]))
            R = fwd[0]
            del fwd, bwd
        f_ms.sort()
        b_ms.sort()
        return dict(fwd_ms=round(f_ms[len(f_ms) // 2], 3), bwd_ms=round(b_ms[len(b_ms) // 2], 3), num_rendered=R)

    out = dict(workload="configs[2]: %d-Gaussian UAV slab, hierarchy cut of %d nodes with interpolation weights, %dx%d, SH 3, "
               "colour only" % (n, P, W, H))
    from hidegs_b200 import _lib
    _lib.profile_enable(True)
    out["ours"] = time_impl(ru.OUR_C, after_warmup=_lib.profile_collect)
    _lib.profile_enable(False)
    st = _lib.profile_collect()
    out["ours"]["stage_ms"] = {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items()}
    if ru.ref_available():
        out["reference"] = time_impl(ru.ref_module())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
