"""Where the HOST time of a training step goes: torch.profiler over two steps of the UAV recipe (CPU activities),
top ops by self CPU time.   python tools/train_cpu_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    from hidegs_b200 import synthetic as syn, trainer as tr
    dev = torch.device("cuda", 0)
    W, H = bench.WIDTH, bench.HEIGHT
    scene = syn.make_uav_scene(2_000_000, seed=0)
    cams = [syn.uav_camera(i, (j + i) % 8, width=W, height=H).to(dev) for i in range(8) for j in range(8)][:8]
    gts = [g.to(dev) for g in bench.make_gt_images(8, dev)]
    params = tr.GaussianParams.from_scene(scene, dev, spatial_order=True)
    t = tr.ViewShardedTrainer(params, torch.zeros(3, device=dev), cache_ground_truth=True, start_iteration=1000)
    views = list(zip(cams, gts))
    for _ in range(3):
        t.step(views)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU], record_shapes=False) as prof:
        for _ in range(2):
            t.step(views)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))


if __name__ == "__main__":
    main()
