"""Generate tests/golden/geometry_ref_*.npz from the reference's OWN code (build container only, needs /root/reference).

  * utils/graphics_utils.py is imported unmodified: normal_from_depth_image on a seeded depth map, plus the gradient of
    a seeded linear functional of  render_normal(depth) * alpha  w.r.t. the depth (torch autograd through the
    reference's ops);
  * scene/gaussian_model.py cannot be imported (plyfile, pytorch3d, simple_knn._C, gaussian_hierarchy._C are not
    installable offline), so the source text of GaussianModel.get_rotation_matrix / get_smallest_axis / get_normal is
    cut out with `ast` and executed as methods of a duck-typed model; pytorch3d's quaternion_to_matrix (the one absent
    symbol) is the published formula restated in oracle/geometry_oracle.py.  The input_all_map lines of render()
    (gaussian_renderer/__init__.py:161-169) are executed the same way;
  * scene/OurAdam.py is imported unmodified and stepped on CPU tensors (dense, index-list and bool-mask `relevant`).
"""
import ast
import importlib.util
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geometry_utils_t as gt  # noqa: E402
from oracle import geometry_oracle as go  # noqa: E402

REF = "/root/reference"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cut_methods(path, cls, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in names:
                    out[f.name] = textwrap.dedent(ast.get_source_segment(src, f))
    return out


def cut_lines(path, first, last):
    return textwrap.dedent("".join(open(path).readlines()[first - 1:last]))




def main():
    gu = load("ref_graphics_utils", os.path.join(REF, "utils/graphics_utils.py"))
    oa = load("ref_our_adam", os.path.join(REF, "scene/OurAdam.py"))
    # torch 2.3 (the reference's pin) had Optimizer._cuda_graph_capture_health_check; torch 2.11 renamed it.  It is a
    # no-op outside CUDA-graph capture, so the missing hook is stubbed on the imported class.
    if not hasattr(oa.Adam, "_cuda_graph_capture_health_check"):
        oa.Adam._cuda_graph_capture_health_check = lambda self: None

    # ---- GaussianModel.get_normal & friends on a duck-typed model
    meth = cut_methods(os.path.join(REF, "scene/gaussian_model.py"), "GaussianModel",
                       ("get_rotation_matrix", "get_smallest_axis", "get_normal"))
    assert len(meth) == 3, meth.keys()
    ns = {"torch": torch, "quaternion_to_matrix": go.quaternion_to_matrix}
    body = "class Model:\n" + "".join(textwrap.indent(m, "    ") + "\n" for m in meth.values())
    exec(body, ns)
    Model = ns["Model"]
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference calls .cuda() on camera tensors; identity on this CPU box

    prologue_src = cut_lines(os.path.join(REF, "gaussian_renderer/__init__.py"), 161, 169)
    prologue_src = prologue_src.replace(".cuda()", "")

    for name, p in gt.GEOMETRY_CASES.items():
        c = gt.make_geometry_inputs(**p)
        out = {}
        # prologue
        pc = Model()
        xyz = c["xyz"].clone().requires_grad_(True)
        rot = c["rotation"].clone().requires_grad_(True)
        pc._xyz = xyz
        pc.get_scaling = c["scaling"]
        pc.get_rotation = rot

        class Cam:
            world_view_transform = c["view"]
            camera_center = c["campos"]
        env = {"torch": torch, "pc": pc, "viewpoint_camera": Cam, "means3D": xyz}
        exec(prologue_src, env)
        am = env["input_all_map"]
        g_xyz, g_rot = torch.autograd.grad((am * c["g_all_map"]).sum(), (xyz, rot))
        out["all_map"], out["all_map_grad_xyz"], out["all_map_grad_rot"] = am.detach().numpy(), g_xyz.numpy(), g_rot.numpy()
        # epilogue: the reference's normal_from_depth_image, permuted as render_normal does, times alpha
        depth = c["depth"].clone().requires_grad_(True)
        K = go.intrinsic_matrix(*c["K"])
        n = gu.normal_from_depth_image(depth, K, torch.eye(4)).permute(2, 0, 1)
        dn = n * c["alpha"][None]
        gd, = torch.autograd.grad((dn * c["g_normal"]).sum(), depth)
        out["render_normal"], out["depth_normal"], out["depth_normal_grad"] = n.detach().numpy(), dn.detach().numpy(), gd.numpy()
        # Adam: 3 steps each of dense / indexed / masked on independent copies
        for mode in ("dense", "index", "mask"):
            prm = torch.nn.Parameter(c["adam_p"].clone())
            opt = oa.Adam([{"params": [prm], "lr": 1.6e-3, "name": "p"}], lr=0.0, eps=1e-15)
            for s in range(3):
                prm.grad = c["adam_g"][s].clone()
                if mode == "dense":
                    rel = torch.empty(0, dtype=torch.long)
                elif mode == "index":
                    rel = c["adam_rel"][s]
                else:
                    rel = torch.zeros(prm.size(0), dtype=torch.bool)
                    rel[c["adam_rel"][s]] = True
                opt.step(rel)
            st = opt.state[prm]
            out["adam_%s_p" % mode] = prm.detach().numpy().copy()
            out["adam_%s_m" % mode] = st["exp_avg"].numpy().copy()
            out["adam_%s_v" % mode] = st["exp_avg_sq"].numpy().copy()
        path = os.path.join(HERE, "geometry_ref_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
