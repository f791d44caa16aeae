"""point_cloud.ply / point_cloud.bin.  Writer side PINNED: the files this package writes are read
back by the reference's own C++ Loader (submodules/gaussianhierarchy/loader.cpp, compiled unmodified into
oracle/_ref/ref_hier_io.so).  Reader side and header text: against the restated oracle (oracle/ply_oracle.py — the
container is written by plyfile==1.1 in the reference, which is not installable here) and through round trips."""
import struct

import numpy as np
import pytest
import torch

from oracle import ply_oracle as po
from hidegs_b200 import ply_io


def _model(n, seed=0, degree=3):
    g = torch.Generator().manual_seed(seed)
    k = (degree + 1) ** 2
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return dict(xyz=r(n, 3) * 10, features_dc=r(n, 1, 3), features_rest=r(n, k - 1, 3) * 0.1, opacity=r(n, 1),
                scaling=r(n, 3) - 3, rotation=r(n, 4))


def _np(m):
    return {k: v.numpy() for k, v in m.items()}


@pytest.mark.parametrize("n", [1, 257])
def test_save_ply_reproduces_the_oracle_bytes(tmp_path, n):
    m = _model(n)
    path = str(tmp_path / "pc" / "point_cloud.ply")  # the directory is created, like mkdir_p in the reference
    ply_io.save_ply(path, **m)
    assert open(path, "rb").read() == po.file_bytes(**_np(m))


@pytest.mark.parametrize("degree", [0, 1, 3])
def test_load_ply_file_matches_the_reference_reader(tmp_path, degree):
    m = _model(100, seed=2, degree=degree)
    path = str(tmp_path / "p.ply")
    open(path, "wb").write(po.file_bytes(**_np(m)))
    got, want = ply_io.load_ply_file(path, degree), po.load_ply_file(path, degree)
    for a, b in zip(got, want):
        assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
    assert got[1].shape == (100, 3, 1) and got[2].shape == (100, 3, (degree + 1) ** 2 - 1) and got[3].shape == (100, 1)


def test_reader_accepts_permuted_columns_extra_fields_comments_and_ascii(tmp_path):
    m = _model(5, seed=3)
    names = po.attribute_names()
    raw = po.file_bytes(**_np(m))
    body = np.frombuffer(raw.split(b"end_header\n", 1)[1], dtype="<f4").reshape(5, len(names))
    perm = list(reversed(range(len(names))))
    hdr = "ply\nformat binary_little_endian 1.0\ncomment made by a test\nelement vertex 5\n"
    hdr += "".join("property float %s\n" % names[i] for i in perm) + "property uchar flag\nelement face 0\nend_header\n"
    rec = np.zeros(5, dtype=[(names[i], "<f4") for i in perm] + [("flag", "u1")])
    for i in perm:
        rec[names[i]] = body[:, i]
    path = str(tmp_path / "perm.ply")
    open(path, "wb").write(hdr.encode() + rec.tobytes())
    ref_path = str(tmp_path / "ref.ply")
    open(ref_path, "wb").write(raw)
    for a, b in zip(ply_io.load_ply_file(path, 3), ply_io.load_ply_file(ref_path, 3)):
        assert np.array_equal(a, b)
    asc = str(tmp_path / "ascii.ply")
    with open(asc, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 5\n" + "".join("property float %s\n" % n for n in names) + "end_header\n")
        for row in body:
            f.write(" ".join(repr(float(x)) for x in row) + "\n")
    for a, b in zip(ply_io.load_ply_file(asc, 3), ply_io.load_ply_file(ref_path, 3)):
        assert np.array_equal(a, b)


def test_round_trip_to_model_layout(tmp_path):
    m = _model(300, seed=4)
    path = str(tmp_path / "rt.ply")
    ply_io.save_ply(path, **m)
    back = ply_io.load_ply(path, degree=3, device="cpu")
    for got, key in zip(back, ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")):
        assert got.shape == m[key].shape and torch.equal(got, m[key]), key


def test_empty_model(tmp_path):
    path = str(tmp_path / "empty.ply")
    ply_io.save_ply(path, **_model(0))
    out = ply_io.load_ply_file(path, 3)
    assert [a.shape[0] for a in out] == [0] * 6 and out[2].shape == (0, 3, 15)


def test_errors(tmp_path):
    bad = str(tmp_path / "bad.ply")
    open(bad, "wb").write(b"not a ply\n")
    with pytest.raises(ValueError, match="not a PLY"):
        ply_io.load_ply_file(bad, 3)
    m = _model(4)
    raw = po.file_bytes(**_np(m))
    open(bad, "wb").write(raw[:-10])
    with pytest.raises(ValueError, match="truncated"):
        ply_io.load_ply_file(bad, 3)
    good = str(tmp_path / "good.ply")
    open(good, "wb").write(raw)
    with pytest.raises(AssertionError):  # wrong SH degree for this file, as the reference asserts (gaussian_model.py:337)
        ply_io.load_ply_file(good, 2)


def test_point_cloud_bin(tmp_path):
    m = _model(17, seed=5)
    path = str(tmp_path / "point_cloud.bin")
    ply_io.save_point_cloud_bin(path, **m)
    raw = open(path, "rb").read()
    assert struct.unpack("i", raw[:4])[0] == 17 and len(raw) == 4 + 17 * 59 * 4
    f = np.frombuffer(raw[4:], dtype="<f4")
    assert np.array_equal(f[:51], m["xyz"].numpy().reshape(-1))
    feats = torch.cat((m["features_dc"], m["features_rest"]), 1).numpy().reshape(-1)
    assert np.array_equal(f[51:51 + 17 * 48], feats)


@pytest.mark.gpu
def test_ply_from_and_to_the_device(cuda_device, tmp_path):
    m = _model(5000, seed=6)
    dev = {k: v.to(cuda_device) for k, v in m.items()}
    path = str(tmp_path / "dev.ply")
    ply_io.save_ply(path, **dev)
    assert open(path, "rb").read() == po.file_bytes(**_np(m))
    back = ply_io.load_ply(path, degree=3, device=cuda_device)
    for got, key in zip(back, ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")):
        assert got.is_cuda and torch.equal(got.cpu(), m[key]), key


def test_done_pt_and_exposure_json(tmp_path):
    m = _model(23, seed=8)
    d = str(tmp_path / "chunk")
    ply_io.save_pt(d, **m)
    import os
    assert sorted(os.listdir(d)) == sorted(list(ply_io._PT_FILES) + ["point_cloud.bin"])
    back = ply_io.load_pt(d)
    for got, key in zip(back, ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")):
        assert torch.equal(got, m[key]), key
    exp = torch.eye(3, 4)[None].repeat(3, 1, 1) + torch.arange(3).view(3, 1, 1) * 0.01
    path = str(tmp_path / "exposure.json")
    ply_io.save_exposures(path, exp, ["a.jpg", "b.jpg", "c.jpg"])
    got = ply_io.load_exposures(path, device="cpu")
    assert list(got) == ["a.jpg", "b.jpg", "c.jpg"] and all(torch.equal(got[n], exp[i]) for i, n in enumerate(got))
    assert ply_io.load_exposures(str(tmp_path / "missing.json"), device="cpu") is None


# ------------------------------------------------------------------ pinned: the reference's own C++ readers
def _ref_loader():
    """submodules/gaussianhierarchy/loader.cpp (Loader::loadPly / loadBin: the hierarchy builder's readers of what
    GaussianModel.save_ply / save_pt write), compiled unmodified into oracle/_ref/ref_hier_io.so by `make -C oracle ref`."""
    import ctypes
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_hier_io.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_hier_io.so not built")
    L = ctypes.CDLL(path)
    if not hasattr(L, "ref_load_ply"):
        pytest.skip("oracle/_ref/ref_hier_io.so predates the point-cloud loader shim")
    return L


def _ref_read(fn, path, skybox=0):
    import ctypes
    n = ctypes.c_int(0)
    null = ctypes.c_void_p(None)
    assert fn(path.encode(), skybox, ctypes.byref(n), null, null, null, null, null, null) == 0
    N = n.value
    out = dict(pos=np.zeros((N, 3), np.float32), shs=np.zeros((N, 16, 3), np.float32), opacity=np.zeros(N, np.float32),
               scale=np.zeros((N, 3), np.float32), rot=np.zeros((N, 4), np.float32), cov=np.zeros((N, 6), np.float32))
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert fn(path.encode(), skybox, ctypes.byref(n), *[ptr(out[k]) for k in ("pos", "shs", "opacity", "scale", "rot", "cov")]) == 0
    return out


@pytest.mark.parametrize("kind", ["ply", "bin"])
@pytest.mark.parametrize("n,skybox", [(1, 0), (257, 0), (300, 40)])
def test_written_files_are_read_back_by_the_reference_loader(tmp_path, kind, n, skybox):
    """PARITY PIN of the writer side of f4: point_cloud.ply / point_cloud.bin written by hidegs_b200.ply_io are parsed by
    the reference's own Loader (header walk, `element vertex N`, the fixed 62-float RichPoint record; loader.cpp:76-160)
    into exactly the model that was saved — positions and SH coefficients bit for bit in the loader's [16][3] order,
    opacity / scale / rotation through its sigmoid / exp / normalisation."""
    L = _ref_loader()
    m = _model(n, seed=11)
    if kind == "ply":
        path = str(tmp_path / "point_cloud.ply")
        ply_io.save_ply(path, **m)
        got = _ref_read(L.ref_load_ply, path, skybox)
    else:
        path = str(tmp_path / "point_cloud.bin")
        ply_io.save_point_cloud_bin(path, **m)
        got = _ref_read(L.ref_load_bin, path, skybox)
    k = slice(skybox, None)
    assert got["pos"].shape[0] == n - skybox
    assert np.array_equal(got["pos"], m["xyz"].numpy()[k])
    want_shs = torch.cat((m["features_dc"], m["features_rest"]), dim=1).numpy()[k]
    assert np.array_equal(got["shs"], want_shs)
    assert np.allclose(got["opacity"], torch.sigmoid(m["opacity"][:, 0]).numpy()[k], rtol=1e-6, atol=1e-7)
    assert np.allclose(got["scale"], torch.exp(m["scaling"]).numpy()[k], rtol=1e-6)
    r = m["rotation"].numpy()[k]
    assert np.allclose(got["rot"], r / np.linalg.norm(r, axis=1, keepdims=True), rtol=1e-5, atol=1e-7)
