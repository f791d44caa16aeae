"""The .hier file format and the static cut: oracle and product against files WRITTEN BY THE
REFERENCE (tests/golden/hier_ref_*.hier, made by tests/golden/make_hier_golden.py through the unmodified
HierarchyWriter / HierarchyLoader / Traversal::expandToTarget)."""
import os

import numpy as np
import pytest
import torch

from oracle import hier_oracle as ho
from hidegs_b200.gaussian_hierarchy import _C as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXP = np.load(os.path.join(GOLD, "hier_ref_expected.npz"))
KEYS = ("pos", "shs", "alphas", "scales", "rot", "nodes", "boxes")


def _inputs():
    return {k: EXP["input_" + k] for k in KEYS}


def _as_oracle_dict(t):
    pos, shs, alpha, scales, rot, nodes, boxes = [x.cpu().numpy() for x in t]
    return dict(pos=pos, shs=shs.reshape(len(pos), 48), alphas=alpha.reshape(-1), scales=scales, rot=rot, nodes=nodes,
                boxes=boxes)


def _same(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype == np.float32:  # bit-exact, NaN payloads aside
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), what
    else:
        assert np.array_equal(a, b), what


# ------------------------------------------------------------------------------------------ oracle vs the reference
@pytest.mark.parametrize("variant", ["f32", "half"])
def test_oracle_load_matches_reference_loader(variant):
    got = ho.load(os.path.join(GOLD, "hier_ref_%s.hier" % variant))
    assert got["compressed"] == (variant == "half")
    for k in KEYS:
        _same(got[k], EXP["%s_%s" % (variant, k)], "%s %s" % (variant, k))


@pytest.mark.parametrize("variant", ["f32", "half"])
def test_oracle_writer_reproduces_reference_bytes(variant):
    g = _inputs()
    mine = ho.file_bytes(g["pos"], g["shs"], g["alphas"], g["scales"], g["rot"], g["nodes"], g["boxes"],
                         compressed=variant == "half")
    theirs = open(os.path.join(GOLD, "hier_ref_%s.hier" % variant), "rb").read()
    assert mine == theirs


@pytest.mark.parametrize("target", [0, 1, 2, 100])
def test_oracle_static_cut_matches_reference(target):
    assert np.array_equal(ho.expand_to_target(EXP["input_nodes"], target), EXP["cut_%d" % target])


# ------------------------------------------------------------------------------------------ product (host C++) vs both
@pytest.mark.parametrize("variant", ["f32", "half"])
def test_load_hierarchy_cpu(variant):
    out = H.load_hierarchy(os.path.join(GOLD, "hier_ref_%s.hier" % variant))
    assert [tuple(t.shape[1:]) for t in out] == [(3,), (16, 3), (1,), (3,), (4,), (7,), (2, 4)]
    assert all(t.device.type == "cpu" for t in out) and out[5].dtype == torch.int32
    got = _as_oracle_dict(out)
    for k in KEYS:
        _same(got[k], EXP["%s_%s" % (variant, k)], "%s %s" % (variant, k))


@pytest.mark.parametrize("variant", ["f32", "half"])
def test_write_hierarchy_cpu_reproduces_reference_bytes(tmp_path, variant):
    g = {k: torch.from_numpy(v) for k, v in _inputs().items()}
    path = str(tmp_path / "out.hier")
    H.write_hierarchy(path, g["pos"], g["shs"].view(-1, 16, 3), g["alphas"].view(-1, 1), g["scales"], g["rot"], g["nodes"],
                      g["boxes"], compressed=variant == "half")
    assert open(path, "rb").read() == open(os.path.join(GOLD, "hier_ref_%s.hier" % variant), "rb").read()


def test_write_defaults_to_the_half_variant(tmp_path):
    g = {k: torch.from_numpy(v) for k, v in _inputs().items()}
    path = str(tmp_path / "out.hier")
    H.write_hierarchy(path, g["pos"], g["shs"], g["alphas"], g["scales"], g["rot"], g["nodes"], g["boxes"])
    assert np.fromfile(path, dtype="<i4", count=1)[0] == -len(g["pos"])  # hierarchy_writer.h:32, hierarchy_writer.cpp:60


@pytest.mark.parametrize("target", [0, 1, 2, 100])
def test_expand_to_target(target):
    got = H.expand_to_target(torch.from_numpy(EXP["input_nodes"]), target)
    assert got.dtype == torch.int32 and np.array_equal(got.numpy(), EXP["cut_%d" % target])


def test_random_round_trips_and_cuts_against_the_oracle(tmp_path):
    for seed in range(4):
        g = ho.synthetic_hierarchy(n_leaves=30 + 50 * seed, seed=10 + seed, branching=(2, 5))
        t = {k: torch.from_numpy(v) for k, v in g.items()}
        for compressed in (False, True):
            path = str(tmp_path / ("r%d_%d.hier" % (seed, compressed)))
            H.write_hierarchy(path, t["pos"], t["shs"], t["alphas"], t["scales"], t["rot"], t["nodes"], t["boxes"],
                              compressed=compressed)
            want = ho.file_bytes(g["pos"], g["shs"], g["alphas"], g["scales"], g["rot"], g["nodes"], g["boxes"],
                                 compressed=compressed)
            assert open(path, "rb").read() == want
            back, ref = _as_oracle_dict(H.load_hierarchy(path)), ho.load(path)
            for k in KEYS:
                _same(back[k], ref[k], k)
        for target in range(0, int(g["nodes"][:, 0].max()) + 2):
            assert np.array_equal(H.expand_to_target(t["nodes"], target).numpy(), ho.expand_to_target(g["nodes"], target))


def test_half_rounding_edge_cases(tmp_path):
    """Every finite half and the values around each rounding boundary go through the writer like numpy's IEEE
    round-to-nearest-even (= half.hpp 2.2's)."""
    halves = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float32)
    mids = (halves[:-1].astype(np.float64) + halves[1:].astype(np.float64)) / 2  # exact ties (representable in fp32)
    vals = np.concatenate([halves, mids.astype(np.float32), np.nextafter(mids.astype(np.float32), np.float32(np.inf)),
                           np.nextafter(mids.astype(np.float32), np.float32(-np.inf)),
                           np.array([65504.0, 65519.99, 65520.0, 1e6, np.inf, 2.0 ** -25, 2.0 ** -26], np.float32)])
    vals = np.concatenate([vals, -vals]).astype(np.float32)
    P = (len(vals) + 47) // 48
    shs = np.zeros((P, 48), np.float32)
    shs.reshape(-1)[:len(vals)] = vals
    z = lambda *s: torch.zeros(*s)  # noqa: E731
    nodes = torch.tensor([[0, -1, 0, P, 0, -1, 0]], dtype=torch.int32)
    path = str(tmp_path / "edge.hier")
    H.write_hierarchy(path, z(P, 3), torch.from_numpy(shs), z(P, 1), z(P, 3), z(P, 4), nodes, z(1, 2, 4), compressed=True)
    with np.errstate(over="ignore"):
        want = shs.astype(np.float16).astype(np.float32)
    got = H.load_hierarchy(path)[1].numpy().reshape(P, 48)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_errors_mirror_the_reference(tmp_path):
    with pytest.raises(RuntimeError, match="File not found!"):  # hierarchy_loader.cpp:38
        H.load_hierarchy(str(tmp_path / "missing.hier"))
    g = {k: torch.from_numpy(v.copy()) for k, v in _inputs().items()}
    g["nodes"][3, 0] = 32001
    with pytest.raises(RuntimeError, match="Would lose information!"):  # hierarchy_writer.cpp:96-97
        H.write_hierarchy(str(tmp_path / "x.hier"), g["pos"], g["shs"], g["alphas"], g["scales"], g["rot"], g["nodes"],
                          g["boxes"], compressed=True)
    with pytest.raises(RuntimeError, match="File not created!"):  # hierarchy_writer.cpp:41
        H.write_hierarchy(str(tmp_path / "no_such_dir" / "x.hier"), g["pos"], g["shs"], g["alphas"], g["scales"], g["rot"],
                          torch.from_numpy(_inputs()["nodes"]), g["boxes"], compressed=False)
    trunc = str(tmp_path / "trunc.hier")
    data = open(os.path.join(GOLD, "hier_ref_half.hier"), "rb").read()
    open(trunc, "wb").write(data[:len(data) // 2])
    with pytest.raises(RuntimeError, match="truncated"):
        H.load_hierarchy(trunc)


def test_empty_hierarchy(tmp_path):
    path = str(tmp_path / "empty.hier")
    z = lambda *s: torch.zeros(*s)  # noqa: E731
    for compressed in (False, True):
        H.write_hierarchy(path, z(0, 3), z(0, 16, 3), z(0, 1), z(0, 3), z(0, 4), torch.zeros(0, 7, dtype=torch.int32),
                          z(0, 2, 4), compressed=compressed)
        out = H.load_hierarchy(path)
        assert [t.shape[0] for t in out] == [0] * 7


# ------------------------------------------------------------------------------------------ device decode / encode
@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["f32", "half"])
def test_load_hierarchy_on_device(cuda_device, variant):
    out = H.load_hierarchy(os.path.join(GOLD, "hier_ref_%s.hier" % variant), device=cuda_device)
    assert all(t.is_cuda for t in out)
    got = _as_oracle_dict(out)
    for k in KEYS:
        _same(got[k], EXP["%s_%s" % (variant, k)], "%s %s" % (variant, k))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["f32", "half"])
def test_write_hierarchy_from_device_reproduces_reference_bytes(cuda_device, tmp_path, variant):
    g = {k: torch.from_numpy(v).to(cuda_device) for k, v in _inputs().items()}
    path = str(tmp_path / "out.hier")
    H.write_hierarchy(path, g["pos"], g["shs"].view(-1, 16, 3), g["alphas"].view(-1, 1), g["scales"], g["rot"], g["nodes"],
                      g["boxes"], compressed=variant == "half")
    assert open(path, "rb").read() == open(os.path.join(GOLD, "hier_ref_%s.hier" % variant), "rb").read()


@pytest.mark.gpu
def test_device_round_trip_large_and_rounding(cuda_device, tmp_path):
    """300k Gaussians through encode -> file -> decode on the device equal the host path bit for bit (both variants);
    a node that would lose information is refused."""
    g = ho.synthetic_hierarchy(n_leaves=2000, seed=5, branching=(2, 6))
    rng = np.random.default_rng(1)
    P = 300_000
    big = dict(pos=rng.normal(0, 50, (P, 3)).astype(np.float32), shs=rng.normal(0, 1, (P, 48)).astype(np.float32),
               alphas=rng.uniform(0, 1, P).astype(np.float32), scales=rng.normal(-3, 2, (P, 3)).astype(np.float32),
               rot=rng.normal(0, 1, (P, 4)).astype(np.float32), nodes=g["nodes"], boxes=g["boxes"])
    big["shs"][:70000].reshape(-1)[:0x7c00 * 2] = np.repeat(np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16), 2)
    big["shs"][1000] *= 1e-7  # subnormal halves
    dev = {k: torch.from_numpy(v).to(cuda_device) for k, v in big.items()}
    cpu = {k: torch.from_numpy(v) for k, v in big.items()}
    for compressed in (False, True):
        a, b = str(tmp_path / "dev.hier"), str(tmp_path / "cpu.hier")
        H.write_hierarchy(a, dev["pos"], dev["shs"], dev["alphas"], dev["scales"], dev["rot"], dev["nodes"], dev["boxes"],
                          compressed=compressed)
        H.write_hierarchy(b, cpu["pos"], cpu["shs"], cpu["alphas"], cpu["scales"], cpu["rot"], cpu["nodes"], cpu["boxes"],
                          compressed=compressed)
        assert open(a, "rb").read() == open(b, "rb").read()
        on_dev, on_cpu = _as_oracle_dict(H.load_hierarchy(a, device=cuda_device)), _as_oracle_dict(H.load_hierarchy(a))
        for k in KEYS:
            _same(on_dev[k], on_cpu[k], k)
    dev["nodes"][7, 3] = 40000
    with pytest.raises(RuntimeError, match="Would lose information!"):
        H.write_hierarchy(str(tmp_path / "bad.hier"), dev["pos"], dev["shs"], dev["alphas"], dev["scales"], dev["rot"],
                          dev["nodes"], dev["boxes"], compressed=True)


def test_host_half_conversion_scalar_and_simd_paths_agree():
    """hier_half.cpp: the F16C path (when the CPU has it) and the scalar fallback give the same bits, and both equal
    numpy's IEEE round-to-nearest-even, on every finite half, every rounding tie and the overflow / underflow edges."""
    import ctypes
    from hidegs_b200 import _lib
    L = _lib.lib()
    halves = np.arange(0, 0x7c01, dtype=np.uint16)
    f = halves.view(np.float16).astype(np.float32)
    mids = ((f[:-2].astype(np.float64) + f[1:-1].astype(np.float64)) / 2).astype(np.float32)
    vals = np.concatenate([f[:-1], mids, np.nextafter(mids, np.float32(np.inf)), np.nextafter(mids, np.float32(-np.inf)),
                           np.array([65504.0, 65519.996, 65520.0, 7e4, 1e30, np.inf, 2.0 ** -25, 2.0 ** -26, 1e-30], np.float32)])
    vals = np.ascontiguousarray(np.concatenate([vals, -vals]).astype(np.float32))
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    for scalar in (0, 1):
        got = np.empty(len(vals), np.uint16)
        L.hg_test_float_to_half(p(vals), p(got), ctypes.c_int64(len(vals)), scalar)
        assert np.array_equal(got, want), ("float->half", scalar)
        back = np.empty(len(halves), np.float32)
        L.hg_test_half_to_float(p(halves), p(back), ctypes.c_int64(len(halves)), scalar)
        assert np.array_equal(back.view(np.uint32), halves.view(np.float16).astype(np.float32).view(np.uint32)), scalar
