"""Generates tests/golden/hier_ref_{f32,half}.hier and hier_ref_expected.npz with the UNMODIFIED reference code
(HierarchyWriter / HierarchyLoader / Traversal::expandToTarget compiled from /root/reference by `make -C oracle ref` into
oracle/_ref/ref_hier_io.so).  Run in the build container:  python tests/golden/make_hier_golden.py"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hier_oracle as ho  # noqa: E402


def ref():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "ref_hier_io.so"))
    lib.ref_hier_load.restype = ctypes.c_int
    lib.ref_hier_write.restype = ctypes.c_int
    lib.ref_expand_to_target.restype = ctypes.c_int
    return lib


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def aligned(a, align=64):
    """Copy into 64-byte aligned storage (the reference casts the pointers to Eigen::Vector4f*, as torch's are)."""
    buf = np.empty(a.nbytes + align, np.uint8)
    off = (-buf.ctypes.data) % align
    out = buf[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


def ref_load(lib, path):
    P, N = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.ref_hier_load(path.encode(), ctypes.byref(P), ctypes.byref(N), *([None] * 7)) == 0
    P, N = P.value, N.value
    out = dict(pos=np.empty((P, 3), np.float32), shs=np.empty((P, 48), np.float32), alphas=np.empty(P, np.float32),
               scales=np.empty((P, 3), np.float32), rot=np.empty((P, 4), np.float32), nodes=np.empty((N, 7), np.int32),
               boxes=np.empty((N, 2, 4), np.float32))
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.ref_hier_load(path.encode(), ctypes.byref(a), ctypes.byref(b), p(out["pos"]), p(out["shs"]),
                             p(out["alphas"]), p(out["scales"]), p(out["rot"]), p(out["nodes"]), p(out["boxes"])) == 0
    return out


def main():
    lib = ref()
    g = ho.synthetic_hierarchy(n_leaves=40, seed=0)
    # exercise the half rounding: ties, subnormals, overflow to inf, negative zero
    g["shs"][0, :8] = [1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 2.0 ** -25, 2.0 ** -24 * 1.5, 65519.9, 65520.0, -0.0, 6.0e-8]
    P, N = len(g["pos"]), len(g["nodes"])
    g = {k: aligned(v) for k, v in g.items()}
    expected = {}
    for name, compressed in (("f32", 0), ("half", 1)):
        path = os.path.join(HERE, "hier_ref_%s.hier" % name)
        assert lib.ref_hier_write(path.encode(), P, N, p(g["pos"]), p(g["shs"]), p(g["alphas"]), p(g["scales"]),
                                  p(g["rot"]), p(g["nodes"]), p(g["boxes"]), compressed) == 0
        back = ref_load(lib, path)
        for k, v in back.items():
            expected["%s_%s" % (name, k)] = v
    for k in ("pos", "shs", "alphas", "scales", "rot", "nodes", "boxes"):
        expected["input_" + k] = g[k]
    for target in (0, 1, 2, 100):
        n = lib.ref_expand_to_target(p(g["nodes"]), target, None, 0)
        out = np.empty(n, np.int32)
        lib.ref_expand_to_target(p(g["nodes"]), target, p(out), n)
        expected["cut_%d" % target] = out
    np.savez_compressed(os.path.join(HERE, "hier_ref_expected.npz"), **expected)
    print("wrote", P, "Gaussians,", N, "nodes")


if __name__ == "__main__":
    main()
