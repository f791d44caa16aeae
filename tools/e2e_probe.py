"""Where does the e2e step spend its time?  (diagnostic, run under gpurun)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hidegs_b200 import synthetic as syn
from hidegs_b200.diff_gaussian_rasterization import GaussianRasterizer, _C

dev = torch.device("cuda:0")
N, W, H = 1_000_000, 1920, 1080
sc = {k: v.to(dev) for k, v in syn.make_scene(N, seed=0).items()}
cam = syn.default_camera(W, H).to(dev)
am = syn.geometry_all_map(sc["means3D"], sc["scales"], sc["rotations"], cam)
params = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacity", "scales", "rotations")}
amp = am.clone().requires_grad_(True)
gt = torch.rand(3, H, W, device=dev)
g = {k: v.to(dev) for k, v in syn.upstream_grads(W, H).items()}


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(6):
    t0 = T()
    rs = syn.raster_settings(cam, dev)
    t1 = T()
    means2D = torch.zeros_like(params["means3D"], requires_grad=True)
    for p in params.values():
        p.grad = None
    out = GaussianRasterizer(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacity"], shs=params["shs"],
                                 scales=params["scales"], rotations=params["rotations"], all_map=amp)
    t2 = T()
    color, radii, obs, amap, pd, inv = out
    loss = (color - gt).abs().mean() + (amap * g["all_map"]).mean() * 1e-3 + (pd * g["plane_depth"]).mean() * 1e-3 + (inv * g["invdepth"]).mean() * 1e-3
    t3 = T()
    loss.backward()
    t4 = T()
    print("it%d settings %.2f ms | fwd(autograd) %.2f | loss %.2f | backward(autograd) %.2f" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))

e_i = torch.empty(0, dtype=torch.int32, device=dev); e_f = torch.empty(0, device=dev)
fa = (torch.zeros(3, device=dev), e_i, e_i, e_f, e_i, sc["means3D"], e_f, am, sc["opacity"], sc["scales"], sc["rotations"], 1.0, e_f,
      cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, sc["shs"], 3, cam.camera_center, False, True, False, True)
for it in range(4):
    t0 = T()
    fwd = _C.rasterize_gaussians(*fa)
    t1 = T()
    R, color, radii, observe, out_all_map, plane_depth, geom, binning, img, invdepth = fwd
    ba = (fa[0], out_all_map, e_i, e_i, e_f, e_i, sc["means3D"], radii, e_f, am, sc["opacity"], sc["scales"], sc["rotations"], 1.0, e_f,
          cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, g["color"], g["all_map"], g["plane_depth"], g["invdepth"],
          sc["shs"], 3, cam.camera_center, geom, R, binning, img, True, False)
    bw = _C.rasterize_gaussians_backward(*ba)
    t2 = T()
    print("direct it%d fwd %.2f ms bwd %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    rs = syn.raster_settings(cam, dev)
    means2D = torch.zeros_like(params["means3D"], requires_grad=True)
    out = GaussianRasterizer(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacity"], shs=params["shs"],
                                 scales=params["scales"], rotations=params["rotations"], all_map=amp)
    loss = (out[0] - gt).abs().mean() + (out[3] * g["all_map"]).mean() * 1e-3
    loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=15))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=15))
