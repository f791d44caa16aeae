"""Replicates bench.py's e2e loop with per-step wall times and allocator statistics."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from hidegs_b200 import synthetic as syn
from hidegs_b200.diff_gaussian_rasterization import GaussianRasterizer

dev = torch.device("cuda:0")
W, H = bench.WIDTH, bench.HEIGHT
scene, _ = bench.load_scene(dev)
bg = torch.zeros(3, device=dev)
g = {k: v.to(dev) for k, v in syn.upstream_grads(W, H, seed=1).items()}
params = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacity", "scales", "rotations")}
gt_host = torch.rand(3, H, W).pin_memory()
w_geo, w_pd, w_inv = g["all_map"] * 1e-3, g["plane_depth"] * 1e-3, g["invdepth"] * 1e-3


def run(vary, n=12, sync_each=True):
    cams = [bench.camera_for(0, s if vary else 0).to(dev) for s in range(n)]
    am = syn.geometry_all_map(scene["means3D"], scene["scales"], scene["rotations"], cams[0]).clone().requires_grad_(True)
    cam_host = [torch.cat([c.world_view_transform.flatten().cpu(), c.full_proj_transform.flatten().cpu(), c.camera_center.cpu()]).pin_memory() for c in cams]
    for s in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st0 = torch.cuda.memory_stats()
        cam = cams[s]
        cd = cam_host[s].to(dev, non_blocking=True)
        gt = gt_host.to(dev, non_blocking=True)
        view, proj, campos = cd[:16].view(4, 4), cd[16:32].view(4, 4), cd[32:35]
        for p in params.values():
            p.grad = None
        rs = syn.raster_settings(cam, dev)._replace(viewmatrix=view, projmatrix=proj, campos=campos, bg=bg)
        means2D = torch.zeros_like(params["means3D"], requires_grad=True)
        t1 = time.perf_counter()
        color, radii, obs, amap, pdepth, inv = GaussianRasterizer(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacity"],
                                                                      shs=params["shs"], scales=params["scales"], rotations=params["rotations"], all_map=am)
        t2 = time.perf_counter()
        loss = (color - gt).abs().mean() + (amap * w_geo).mean() + (pdepth * w_pd).mean() + (inv * w_inv).mean()
        t3 = time.perf_counter()
        loss.backward()
        t4 = time.perf_counter()
        v = float(loss.item())
        t5 = time.perf_counter()
        st1 = torch.cuda.memory_stats()
        print("vary=%d step %2d total %.2f ms: prep %.2f fwd %.2f loss %.2f bwd-launch %.2f item %.2f | cudaMalloc +%d retries +%d reserved %.2f GB" % (
            vary, s, (t5 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3,
            st1["num_device_alloc"] - st0["num_device_alloc"], st1["num_alloc_retries"] - st0["num_alloc_retries"],
            st1["reserved_bytes.all.current"] / 2**30))


run(False)
run(True)
