"""Generate tests/golden/raster_ref_*.npz from the UNMODIFIED reference rasterizer.

Run on a B200 box (`gpurun -- python tests/golden/make_raster_golden.py`) after
`make -C oracle ref` built oracle/_ref/ref_rasterizer_C.so from the reference's
own sources.  Each fixture stores the seeded case parameters (the inputs are
re-generated from them by tests/raster_utils.build_case) and the reference's
outputs: keys, sorted list, tile ranges, n_contrib, final_T, images and all
nine gradient tensors.  Files are written to gpurun_out/golden/ (which gpurun
brings back) and then copied to tests/golden/.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raster_utils as ru  # noqa: E402
from hidegs_b200 import synthetic as syn  # noqa: E402

CASES = {
    "plain": dict(n=1500, W=96, H=64, seed=11, with_hier=False, with_indices=False, render_geo=True, do_depth=True),
    "hier": dict(n=1500, W=96, H=64, seed=12, with_hier=True, with_indices=False, render_geo=True, do_depth=True),
    "raw_indices": dict(n=1500, W=96, H=64, seed=13, with_hier=True, with_indices=True, render_geo=True, do_depth=True),
    "raw_indices_d0": dict(n=1500, W=96, H=64, seed=15, with_hier=True, with_indices=True, render_geo=True, do_depth=True, sh_degree=0),
    "nogeo": dict(n=1500, W=100, H=70, seed=14, with_hier=False, with_indices=False, render_geo=False, do_depth=False),
}


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    REF = ru.ref_module()
    for name, p in CASES.items():
        case = ru.build_case(p["n"], p["W"], p["H"], seed=p["seed"], with_hier=p["with_hier"], with_indices=p["with_indices"],
                             sh_degree=p.get("sh_degree", 3))
        fa = ru.op_args(case, dev, render_geo=p["render_geo"], do_depth=p["do_depth"])
        ref = REF.rasterize_gaussians(*fa)
        torch.cuda.synchronize()
        st = ru.ref_state(ref, case["P"], p["W"], p["H"])
        grads = syn.upstream_grads(p["W"], p["H"], do_depth=p["do_depth"])
        gb = REF.rasterize_gaussians_backward(*ru.bwd_args(fa, ref, grads, dev))
        torch.cuda.synchronize()
        R = ref[0]
        d = dict(params=np.frombuffer(json.dumps(p).encode(), dtype=np.uint8), num_rendered=np.int64(R),
                 color=ref[1].cpu().numpy(), radii=ref[2].cpu().numpy(), out_observe=ref[3].cpu().numpy(),
                 all_map=ref[4].cpu().numpy(), plane_depth=ref[5].cpu().numpy(), invdepth=ref[9].cpu().numpy(),
                 final_T=st["final_T"].cpu().numpy(), n_contrib=st["n_contrib"].cpu().numpy(),
                 ranges=st["ranges"].cpu().numpy(), depths=st["depths"].cpu().numpy(),
                 tiles_touched=st["tiles_touched"].cpu().numpy(), means2D=st["means2D"].cpu().numpy(),
                 conic_opacity=st["conic_opacity"].cpu().numpy(), rgb=st["rgb"].cpu().numpy())
        if R > 0:
            d.update(keys=st["keys"].cpu().numpy(), point_list=st["point_list"].cpu().numpy(),
                     keys_unsorted=st["keys_unsorted"].cpu().numpy())
        for gname, g in zip(ru.GRAD_NAMES, gb):
            d[gname] = g.cpu().numpy()
        path = os.path.join(out_dir, "raster_ref_%s.npz" % name)
        np.savez_compressed(path, **d)
        print(name, "R=%d" % R, "visible=%d" % int((ref[2] > 0).sum()), "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
