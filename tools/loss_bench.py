#!/usr/bin/env python
"""Times the loss path of BASELINE.json configs[0] on the GPU and, beside it, the CPU
restatement of the reference's PyTorch losses on the host cores.

    python tools/loss_bench.py [--iters K] [--warmup W] [--no-cpu] [--noise 0.002|0.05] [--json out.json]

One iteration = forward + backward of
    (1 - 0.2) * l1_loss + 0.2 * (1 - ssim) + frequency_regularization_pyramid_scale(...)      (iteration 2000)
on a seeded 3x1080x1920 render / ground-truth pair with a 100k-Gaussian scaling shim, i.e. every image-space loss
kernel of the training step.  Timed with CUDA events on the current stream; per-function times are taken in a second
pass.  Importable: `measure(dev, iters, warmup, cpu)` returns the dict bench.py embeds as its "losses" object.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

H, W, N_SHIM = 1080, 1920, 100_000


class _Shim:
    def __init__(self, scaling):
        self.get_scaling = scaling


def make_inputs(noise=0.002, seed=0):
    torch.manual_seed(seed)
    gt = F.avg_pool2d(torch.rand(3, H, W)[None], 5, stride=1, padding=2)[0].clamp(0, 1).contiguous()
    render = (gt + noise * torch.randn(3, H, W)).clamp(0, 1).contiguous()
    scaling = torch.rand(N_SHIM, 3) * 0.05
    vis = torch.arange(0, N_SHIM, 2)
    return gt, render, scaling, vis


def _events(fn, iters, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def measure(dev, iters=20, warmup=5, cpu=True, noise=0.002, cpu_iters=2):
    from hidegs_b200 import loss_utils as lu
    from hidegs_b200.frequency_regularization import frequency_regularization_pyramid_scale as freg
    gt_c, render_c, scaling_c, vis_c = make_inputs(noise)
    gt, vis = gt_c.to(dev), vis_c.to(dev)
    render = render_c.to(dev).requires_grad_(True)
    scaling = scaling_c.to(dev).requires_grad_(True)
    out = {}

    def full():
        render.grad = None
        scaling.grad = None
        l1 = lu.l1_loss(render, gt)
        s = lu.ssim(render, gt)
        fr, _m, _i = freg(render, gt, _Shim(scaling), None, None, vis, 2000)
        loss = 0.8 * l1 + 0.2 * (1.0 - s) + fr
        loss.backward()
        return loss

    def only(f):
        def run():
            render.grad = None
            scaling.grad = None
            f().backward()
        return run

    out["gpu_ms"] = round(_events(full, iters, warmup), 4)
    from hidegs_b200 import _lib
    torch.cuda.synchronize()
    _lib.lib().hg_reset_launch_count()
    full()
    out["gpu_launches"] = int(_lib.lib().hg_launch_count())  # this library's kernels in one forward + backward
    out["gpu_loss_value"] = float(full().item())
    out["gpu_ms_parts"] = {
        "l1": round(_events(only(lambda: lu.l1_loss(render, gt)), iters, 2), 4),
        "ssim": round(_events(only(lambda: lu.ssim(render, gt)), iters, 2), 4),
        "frequency_regularization": round(_events(only(lambda: freg(render, gt, _Shim(scaling), None, None, vis, 2000)[0]), iters, 2), 4),
        "img_grad_weight": round(_events(lambda: lu.get_img_grad_weight(gt), iters, 2), 4),
    }
    # algorithmic bytes of the loss path (SURVEY.md §8(d) "Loss-path bytes": 115 B per level-0 pixel)
    out["algorithmic_bytes"] = 115 * H * W
    out["achieved_gbs"] = round(out["algorithmic_bytes"] / (out["gpu_ms"] * 1e-3) / 1e9, 1)
    out["workload"] = "configs[0]: L1 + SSIM + frequency_regularization_pyramid_scale fwd+bwd, 3x1080x1920, noise %.3g" % noise
    if cpu:
        import oracle.loss_oracle as lo
        torch.set_num_threads(os.cpu_count() or 1)
        ts = []
        val = None
        for _ in range(cpu_iters):
            r = render_c.clone().requires_grad_(True)
            sc = scaling_c.clone().requires_grad_(True)
            t0 = time.perf_counter()
            l1 = lo.l1_loss(r, gt_c)
            s = lo.ssim(r, gt_c)
            fr, _m, _i = lo.frequency_regularization_pyramid_scale(r, gt_c, _Shim(sc), None, None, vis_c, 2000)
            loss = 0.8 * l1 + 0.2 * (1.0 - s) + fr
            loss.backward()
            ts.append(time.perf_counter() - t0)
            val = float(loss.item())
        out["cpu_ms"] = round(min(ts) * 1e3, 1)
        out["cpu_cores"] = os.cpu_count() or 1
        out["cpu_kind"] = "port (oracle/loss_oracle.py: torch-CPU restatement of the reference's PyTorch losses), %d iterations, best" % cpu_iters
        out["cpu_loss_value"] = val
        out["speedup_vs_cpu"] = round(out["cpu_ms"] / out["gpu_ms"], 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--noise", type=float, default=0.002)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("loss_bench.py needs a CUDA device (the loss path has no CPU fallback)")
    dev = torch.device("cuda", 0)
    res = measure(dev, a.iters, a.warmup, not a.no_cpu, a.noise)
    s = json.dumps(res)
    print(s, flush=True)
    if a.json:
        with open(a.json, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
