#!/usr/bin/env python
"""The bench's training leg alone (configs[4] recipe by default), for an ncu launch list of exactly what it times:
    ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 900 --csv --log-file out.csv \
        python tools/train_leg_probe.py [--recipe uav|c2] [--views 8] [--steps 2]"""
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
    ap.add_argument("--recipe", default="uav")
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--n", type=int, default=2_000_000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = bench.train_leg(dev, 0, 1, False, steps=a.steps, warmup=3, views_per_rank=a.views, n_gauss=a.n, recipe=a.recipe)
    print(json.dumps({k: out[k] for k in ("views_per_s", "ms_per_step", "ms_per_view", "gpu_launches", "phases")}))


if __name__ == "__main__":
    main()
