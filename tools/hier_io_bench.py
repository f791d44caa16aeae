#!/usr/bin/env python
"""Host-side measurement of the .hier path: hg_hier_load / hg_hier_write against the UNMODIFIED
reference loader / writer (oracle/_ref/ref_hier_io.so, built by `make -C oracle ref`) on the same synthetic file.
CPU only (the device decode path is covered by tests/test_hierarchy_io.py on the GPU box).

    python tools/hier_io_bench.py [--gaussians 1000000] [--out /tmp]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hidegs_b200.gaussian_hierarchy import _C as H  # noqa: E402


def aligned(a, align=64):
    buf = np.empty(a.nbytes + align, np.uint8)
    off = (-buf.ctypes.data) % align
    out = buf[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


def best(fn, n=3):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--out", default="/tmp")
    a = ap.parse_args()
    P = a.gaussians
    N = P // 2
    rng = np.random.default_rng(0)
    g = dict(pos=rng.normal(0, 50, (P, 3)).astype(np.float32), shs=rng.normal(0, 1, (P, 48)).astype(np.float32),
             alphas=rng.uniform(0, 1, P).astype(np.float32), scales=rng.normal(-3, 1, (P, 3)).astype(np.float32),
             rot=rng.normal(0, 1, (P, 4)).astype(np.float32),
             nodes=rng.integers(0, 1000, (N, 7)).astype(np.int32), boxes=rng.normal(0, 50, (N, 2, 4)).astype(np.float32))
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    ga = {k: aligned(v) for k, v in g.items()}
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    ref_path = os.path.join(ROOT, "oracle", "_ref", "ref_hier_io.so")
    ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
    out = {"gaussians": P, "nodes": N}
    for name, comp in (("f32", False), ("half", True)):
        ours_path, ref_file = os.path.join(a.out, "ours_%s.hier" % name), os.path.join(a.out, "ref_%s.hier" % name)
        tw = best(lambda: H.write_hierarchy(ours_path, t["pos"], t["shs"], t["alphas"], t["scales"], t["rot"], t["nodes"],
                                            t["boxes"], compressed=comp))
        tl = best(lambda: H.load_hierarchy(ours_path))
        size = os.path.getsize(ours_path)
        out[name] = {"file_MB": round(size / 1e6, 1), "ours_write_s": round(tw, 3), "ours_load_s": round(tl, 3),
                     "ours_load_GBps": round(size / tl / 1e9, 2)}
        if ref is not None:
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)  # the reference writer prints to stdout
            try:
                trw = best(lambda: ref.ref_hier_write(ref_file.encode(), P, N, p(ga["pos"]), p(ga["shs"]), p(ga["alphas"]),
                                                      p(ga["scales"]), p(ga["rot"]), p(ga["nodes"]), p(ga["boxes"]),
                                                      int(comp)))
            finally:
                os.dup2(saved, 1)
                os.close(devnull)
            assert open(ref_file, "rb").read() == open(ours_path, "rb").read(), "file bytes differ"
            bufs = dict(pos=np.empty((P, 3), np.float32), shs=np.empty((P, 48), np.float32), alphas=np.empty(P, np.float32),
                        scales=np.empty((P, 3), np.float32), rot=np.empty((P, 4), np.float32),
                        nodes=np.empty((N, 7), np.int32), boxes=np.empty((N, 2, 4), np.float32))
            a_, b_ = ctypes.c_int(0), ctypes.c_int(0)
            trl = best(lambda: ref.ref_hier_load(ours_path.encode(), ctypes.byref(a_), ctypes.byref(b_), p(bufs["pos"]),
                                                 p(bufs["shs"]), p(bufs["alphas"]), p(bufs["scales"]), p(bufs["rot"]),
                                                 p(bufs["nodes"]), p(bufs["boxes"])))
            out[name].update({"ref_write_s": round(trw, 3), "ref_load_s": round(trl, 3), "files_identical": True,
                              "load_speedup": round(trl / tl, 2), "write_speedup": round(trw / tw, 2)})
        for f in (ours_path, ref_file):
            if os.path.exists(f):
                os.remove(f)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
