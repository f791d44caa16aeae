"""CPU: oracle/geometry_oracle.py against the fixtures produced by the reference's own code
(tests/golden/make_geometry_golden.py) — prologue (get_normal / input_all_map) with gradients, depth -> normal with
gradients, OurAdam steps — plus self-consistency of the brute-force 3-NN."""
import os

import numpy as np
import pytest
import torch

import geometry_utils_t as gt
from oracle import geometry_oracle as go

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("name", sorted(gt.GEOMETRY_CASES))
def test_prologue_matches_reference(name):
    c = gt.make_geometry_inputs(**gt.GEOMETRY_CASES[name])
    ref = np.load(os.path.join(GOLD, "geometry_ref_%s.npz" % name))
    xyz = c["xyz"].clone().requires_grad_(True)
    rot = c["rotation"].clone().requires_grad_(True)
    am = go.input_all_map(xyz, c["scaling"], rot, c["view"], c["campos"])
    g_xyz, g_rot = torch.autograd.grad((am * c["g_all_map"]).sum(), (xyz, rot))
    assert np.array_equal(am.detach().numpy(), ref["all_map"])
    assert rel(g_xyz.numpy(), ref["all_map_grad_xyz"]) < 1e-6 and rel(g_rot.numpy(), ref["all_map_grad_rot"]) < 1e-6


@pytest.mark.parametrize("name", sorted(gt.GEOMETRY_CASES))
def test_depth_normal_matches_reference(name):
    c = gt.make_geometry_inputs(**gt.GEOMETRY_CASES[name])
    ref = np.load(os.path.join(GOLD, "geometry_ref_%s.npz" % name))
    depth = c["depth"].clone().requires_grad_(True)
    K = go.intrinsic_matrix(*c["K"])
    n = go.render_normal(depth, K)
    dn = go.depth_normal(depth[None], c["alpha"][None], K)
    gd, = torch.autograd.grad((dn * c["g_normal"]).sum(), depth)
    assert np.array_equal(n.detach().numpy(), ref["render_normal"])
    assert np.array_equal(dn.detach().numpy(), ref["depth_normal"])
    assert rel(gd.numpy(), ref["depth_normal_grad"]) < 1e-6
    # border is zero, interior has unit length
    nn = n.detach().numpy()
    assert not nn[:, 0].any() and not nn[:, -1].any() and not nn[:, :, 0].any() and not nn[:, :, -1].any()
    assert np.allclose(np.linalg.norm(nn[:, 1:-1, 1:-1], axis=0), 1.0, atol=1e-5)


@pytest.mark.parametrize("mode", ["dense", "index", "mask"])
def test_adam_matches_reference(mode):
    c = gt.make_geometry_inputs(**gt.GEOMETRY_CASES["small"])
    ref = np.load(os.path.join(GOLD, "geometry_ref_small.npz"))
    p = c["adam_p"].clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s in range(3):
        relv = None if mode == "dense" else c["adam_rel"][s]
        go.adam_step(p, c["adam_g"][s], m, v, s + 1, 1.6e-3, eps=1e-15, relevant=relv)
    assert rel(p.numpy(), ref["adam_%s_p" % mode]) < 1e-6
    assert rel(m.numpy(), ref["adam_%s_m" % mode]) < 1e-6 and rel(v.numpy(), ref["adam_%s_v" % mode]) < 1e-6
    if mode != "dense":  # untouched rows keep parameter and state
        untouched = np.setdiff1d(np.arange(p.size(0)), np.concatenate([r.numpy() for r in c["adam_rel"]]))
        assert np.array_equal(p.numpy()[untouched], c["adam_p"].numpy()[untouched]) and not m.numpy()[untouched].any()


def test_normal_consistency_loss_definition():
    c = gt.make_geometry_inputs(**gt.GEOMETRY_CASES["small"])
    K = go.intrinsic_matrix(*c["K"])
    am = c["out_all_map"]
    loss = go.normal_consistency_loss(c["depth"][None], am, K, c["image_weight"], 0.015)
    dn = go.depth_normal(c["depth"][None], am[3:4], K)
    manual = 0.015 * (c["image_weight"] * (dn - am[0:3]).abs().sum(0)).mean()
    assert abs(loss.item() - manual.item()) <= 1e-7 * abs(manual.item())


def test_knn_bruteforce_properties():
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(500, 3, generator=g).numpy()
    d = go.dist_knn3(pts, chunk=128)
    # against a float64 brute force
    p64 = pts.astype(np.float64)
    D = ((p64[:, None] - p64[None]) ** 2).sum(-1)
    np.fill_diagonal(D, np.inf)
    want = np.sort(D, axis=1)[:, :3].mean(1)
    assert np.allclose(d, want, rtol=1e-5)
    # a duplicated point counts as a neighbour at distance 0
    pts2 = np.concatenate([pts, pts[:1]])
    d2 = go.dist_knn3(pts2)
    assert d2[0] <= d[0] and d2[-1] == d2[0]
    # fewer than 4 points: missing neighbours stay at FLT_MAX as in the reference
    assert not np.isfinite(go.dist_knn3(pts[:2])).all() or go.dist_knn3(pts[:2])[0] > 1e37
