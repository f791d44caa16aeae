"""First-contact GPU diagnostic: ours vs the reference CUDA build vs the CPU oracle.
Prints mismatch counts / error magnitudes instead of asserting (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raster_utils as ru  # noqa: E402
from hidegs_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
REF = ru.ref_module() if ru.ref_available() else None
print("reference .so:", REF is not None)


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(True), torch.cuda.Event(True)
    ts = []
    for _ in range(it):
        s[0].record(); fn(); s[1].record(); torch.cuda.synchronize()
        ts.append(s[0].elapsed_time(s[1]))
    return float(np.median(ts)), float(np.min(ts))


def compare(n, W, H, seed=0, with_hier=False, with_indices=False, render_geo=True, do_depth=True, oracle=True,
            timing=False, tag=""):
    print("\n==== case %s n=%d %dx%d hier=%s idx=%s geo=%s depth=%s" % (tag, n, W, H, with_hier, with_indices, render_geo, do_depth))
    case = ru.build_case(n, W, H, seed=seed, with_hier=with_hier, with_indices=with_indices)
    P = case["P"]
    fa = ru.op_args(case, dev, render_geo=render_geo, do_depth=do_depth)
    ours = ru.OUR_C.rasterize_gaussians(*fa)
    torch.cuda.synchronize()
    so = ru.our_state(ours, P, W, H)
    R = ours[0]
    vis = int((ours[2] > 0).sum())
    print("ours: R=%d visible=%d sum n_contrib=%d" % (R, vis, int(so["n_contrib"].long().sum())))
    grads = syn.upstream_grads(W, H, do_depth=do_depth)
    ours_b = ru.OUR_C.rasterize_gaussians_backward(*ru.bwd_args(fa, ours, grads, dev))
    torch.cuda.synchronize()
    if REF is not None:
        ref = REF.rasterize_gaussians(*fa)
        torch.cuda.synchronize()
        sr = ru.ref_state(ref, P, W, H)
        print("ref : R=%d" % ref[0])
        if ref[0] == R and R > 0:
            print("  keys_unsorted mismatches:", int((so["keys_unsorted"] != sr["keys_unsorted"]).sum()))
            print("  keys (sorted) mismatches:", int((so["keys"] != sr["keys"]).sum()))
            print("  point_list mismatches   :", int((so["point_list"] != sr["point_list"]).sum()))
            print("  ranges mismatches       :", int((so["ranges"] != sr["ranges"]).sum()))
        m = ours[2] > 0
        print("  radii mismatches        :", int((ours[2] != ref[2]).sum()))
        print("  depth bit mismatches    :", int((so["depths"][m].view(torch.int32) != sr["depths"][m].view(torch.int32)).sum()))
        print("  tiles_touched mismatches:", int((so["tiles_touched"] != sr["tiles_touched"]).sum()))
        rec = so["records"]
        print("  means2D bit mismatches  :", int((rec[m][:, 0:2].contiguous().view(torch.int32) != sr["means2D"][m].contiguous().view(torch.int32)).sum()))
        co = torch.stack([rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5]], 1)
        print("  conic_opacity bit mism. :", int((co[m].contiguous().view(torch.int32) != sr["conic_opacity"][m].contiguous().view(torch.int32)).sum()))
        if fa[12].numel() == 0:
            print("  cov3D bit mismatches    :", int((so["cov3D"][m].contiguous().view(torch.int32) != sr["cov3D"][m].contiguous().view(torch.int32)).sum()))
        rgb = rec[:, 6:9]
        print("  rgb max abs err         : %.3g" % float((rgb[m] - sr["rgb"][m]).abs().max()))
        print("  n_contrib mismatches    :", int((so["n_contrib"] != sr["n_contrib"]).sum()))
        print("  final_T bit mismatches  :", int((so["final_T"].view(torch.int32) != sr["final_T"].view(torch.int32)).sum()))
        print("  out_observe mismatches  :", int((ours[3] != ref[3]).sum()))
        for name, i in (("color", 1), ("all_map", 4), ("plane_depth", 5), ("invdepth", 9)):
            a, b = ours[i], ref[i]
            if a.numel():
                fin = torch.isfinite(b) & torch.isfinite(a)
                print("  %-12s max abs err %.3g (max |ref| %.3g) bit mismatches %d, nonfinite ours/ref %d/%d" % (
                    name, float((a[fin] - b[fin]).abs().max()), float(b[fin].abs().max()),
                    int((a.view(torch.int32) != b.view(torch.int32)).sum()), int((~torch.isfinite(a)).sum()), int((~torch.isfinite(b)).sum())))
        ref_b = REF.rasterize_gaussians_backward(*ru.bwd_args(fa, ref, grads, dev))
        torch.cuda.synchronize()
        for name, a, b in zip(ru.GRAD_NAMES, ours_b, ref_b):
            print("  %-14s max err %.3g scale %.3g frac>1e-3rel %.3g" % ((name,) + ru.rel_report(a, b)))
    if oracle:
        o = ru.oracle_for_case(case, render_geo=render_geo, do_depth=do_depth, nthreads=os.cpu_count())
        t = time.time(); oo = o.forward(); tf = time.time() - t
        print("oracle: R=%d (%.2fs)" % (oo["num_rendered"], tf))
        if oo["num_rendered"] == R and R > 0:
            print("  keys mismatches vs oracle      :", int((so["keys"].cpu().numpy().view(np.uint64) != oo["keys"]).sum()))
            print("  point_list mismatches vs oracle:", int((so["point_list"].cpu().numpy().view(np.uint32) != oo["point_list"]).sum()))
            print("  ranges mismatches vs oracle    :", int((so["ranges"].cpu().numpy().view(np.uint32) != oo["ranges"]).sum()))
        print("  radii mismatches vs oracle     :", int((ours[2].cpu().numpy() != oo["radii"]).sum()))
        print("  n_contrib mismatches vs oracle :", int((so["n_contrib"].cpu().numpy().view(np.uint32) != oo["n_contrib"]).sum()))
        print("  out_observe mismatches         :", int((ours[3].cpu().numpy() != oo["out_observe"]).sum()))
        for name, i, k in (("color", 1, "color"), ("all_map", 4, "all_map"), ("plane_depth", 5, "plane_depth"), ("invdepth", 9, "invdepth")):
            a, b = ours[i].cpu().numpy(), oo[k]
            if a.size:
                d = np.abs(a - b); fin = np.isfinite(d)
                print("  %-12s max abs err vs oracle %.3g, #>1e-4: %d" % (name, float(d[fin].max()), int((d[fin] > 1e-4).sum())))
        t = time.time()
        og = o.backward(grads["color"].numpy(), grads["all_map"].numpy(), grads["plane_depth"].numpy(),
                        grads["invdepth"].numpy() if do_depth else None)
        print("  oracle backward %.2fs" % (time.time() - t))
        for name, a in zip(ru.GRAD_NAMES, ours_b):
            b = torch.from_numpy(og[name])
            print("  %-14s vs oracle: max err %.3g scale %.3g frac>1e-3rel %.3g" % ((name,) + ru.rel_report(a.cpu(), b)))
    if timing:
        fo = lambda: ru.OUR_C.rasterize_gaussians(*fa)  # noqa: E731
        bo = lambda: ru.OUR_C.rasterize_gaussians_backward(*ru.bwd_args(fa, ours, grads, dev))  # noqa: E731
        tf, tb = timeit(fo), timeit(bo)
        print("TIMING ours fwd %.3f ms (min %.3f) bwd %.3f ms (min %.3f) -> %.1f Mpix/s" % (tf[0], tf[1], tb[0], tb[1], W * H / (tf[0] + tb[0]) / 1e3))
        if REF is not None:
            fr = lambda: REF.rasterize_gaussians(*fa)  # noqa: E731
            br = lambda: REF.rasterize_gaussians_backward(*ru.bwd_args(fa, ref, grads, dev))  # noqa: E731
            tf2, tb2 = timeit(fr), timeit(br)
            print("TIMING ref  fwd %.3f ms (min %.3f) bwd %.3f ms (min %.3f) -> %.1f Mpix/s" % (tf2[0], tf2[1], tb2[0], tb2[1], W * H / (tf2[0] + tb2[0]) / 1e3))


if __name__ == "__main__":
    compare(3000, 160, 96, tag="tiny")
    compare(20000, 320, 192, seed=1, tag="small")
    compare(20000, 320, 192, seed=2, with_hier=True, tag="hier-ts-kids")
    compare(20000, 320, 192, seed=3, with_hier=True, with_indices=True, tag="raw-indices-parents")
    compare(20000, 320, 192, seed=4, render_geo=False, do_depth=False, tag="nogeo-nodepth")
    compare(200000, 960, 540, seed=5, oracle=True, timing=True, tag="mid")
    compare(1000000, 1920, 1080, seed=0, oracle=False, timing=True, tag="config2")
