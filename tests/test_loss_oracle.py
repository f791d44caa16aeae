"""Pins oracle/loss_oracle.py against golden vectors produced by the reference's own loss code
(tests/golden/loss_ref_*.npz, made by tests/golden/make_loss_golden.py) and against the known
answers of BASELINE.md §3 (config 1, 3x1080x1920)."""
import glob
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import loss_utils_t as lt
from oracle import loss_oracle as lo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loss_ref_*.npz")))


def close(a, b, rtol=1e-4, atol=1e-7):
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_loss_goldens_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN)
def test_oracle_matches_reference_losses(path):
    gold = np.load(path)
    p = json.loads(bytes(gold["params"]).decode())
    inp = lt.make_loss_inputs(**p)
    r = inp["render"].clone().requires_grad_(True)
    gt = inp["gt"]
    for name, fn in (("l1", lo.l1_loss), ("l2", lo.l2_loss), ("ssim", lo.ssim)):
        v = fn(r, gt)
        g, = torch.autograd.grad(v, r)
        assert close(v.item(), gold[name]), name
        assert close(g.numpy(), gold[name + "_grad"], rtol=1e-3, atol=1e-9), name
    assert close(lo.ssim(torch.stack([r.detach(), gt]), torch.stack([gt, gt]), size_average=False).numpy(), gold["ssim_batched"])
    assert close(lo.get_img_grad_weight(gt).numpy(), gold["grad_weight"], atol=1e-6)
    a = inp["patch_ref"].clone().requires_grad_(True)
    b = inp["patch_nea"].clone().requires_grad_(True)
    ncc, mask = lo.lncc(a, b)
    ga, gb = torch.autograd.grad((ncc * inp["patch_w"]).sum(), (a, b))
    assert close(ncc.detach().numpy(), gold["lncc"], atol=1e-6) and np.array_equal(mask.numpy(), gold["lncc_mask"])
    assert close(ga.numpy(), gold["lncc_grad_ref"], rtol=1e-3, atol=1e-6) and close(gb.numpy(), gold["lncc_grad_nea"], rtol=1e-3, atol=1e-6)
    scal = inp["scaling"].clone().requires_grad_(True)
    total, hf, info = lo.frequency_regularization_pyramid_scale(r, gt, lt.GaussiansShim(scal), None, None, inp["visibility"], 2000)
    gr, gs = torch.autograd.grad(total, (r, scal))
    assert close(total.item(), gold["freq_total"], rtol=1e-3)
    assert close(info["freq_loss"], gold["info_freq_loss"], rtol=1e-3) and close(info["scale_loss"], gold["info_scale_loss"], rtol=1e-4)
    assert info["high_freq_pixels"] == float(gold["info_high_freq_pixels"])
    assert np.array_equal(hf.numpy().astype(np.uint8), gold["freq_mask"])
    assert close(info["freq_band_energies"], gold["info_band_energies"], rtol=1e-4)
    for lvl in range(3):
        assert close(info["levels"][lvl]["spatial"], gold["lvl%d_spatial" % lvl], rtol=1e-4)
        assert close(info["levels"][lvl]["fft"], gold["lvl%d_fft" % lvl], rtol=1e-3)
    scale = np.abs(gold["freq_grad_render"]).max()
    assert np.abs(gr.numpy() - gold["freq_grad_render"]).max() <= 1e-3 * scale + 1e-12
    assert close(gs.numpy(), gold["freq_grad_scaling"], rtol=1e-4, atol=1e-12)


def test_warmup_gate_and_signature():
    z, m, info = lo.frequency_regularization_pyramid_scale(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8), None, None, None,
                                                          torch.zeros(0), iteration=10)
    assert float(z) == 0.0 and m is None and info == {"warmup": True}


def test_oracle_known_answers_config1():
    """reference values measured during the survey (far variant, seed 0, 3x1080x1920)."""
    torch.manual_seed(0)
    gt = F.avg_pool2d(torch.rand(3, 1080, 1920)[None], 5, stride=1, padding=2)[0].clamp(0, 1)
    render = (gt + 0.05 * torch.randn(3, 1080, 1920)).clamp(0, 1)
    scaling = torch.rand(100000, 3) * 0.05
    vis = torch.arange(0, 100000, 2)
    total, hf, info = lo.frequency_regularization_pyramid_scale(render, gt, lt.GaussiansShim(scaling), None, None, vis, 2000)
    assert close(total.item(), 2.0789e-05, rtol=2e-4)
    assert close(info["freq_loss"], 0.016487, rtol=2e-4) and close(info["scale_loss"], 8.6037e-04, rtol=2e-4)
    assert info["high_freq_pixels"] == 32364
    assert close(info["freq_band_energies"], [197.62, 74.72, 28.48, 14.67], rtol=5e-4)  # BASELINE.md rounds to 2 decimals
    lv = info["levels"]
    assert close(lv[0]["spatial"], 0.019005, rtol=2e-4) and lv[0]["fft"] == 10.0 and close(lv[0]["level"], 0.1)
    assert close(lv[1]["spatial"], 0.004740, rtol=2e-4) and close(lv[1]["fft"], 1.7024, rtol=2e-4)
    assert close(lv[2]["spatial"], 0.001185, rtol=5e-4) and close(lv[2]["fft"], 0.19548, rtol=2e-4) and close(lv[2]["level"], 0.05947, rtol=5e-4)
    assert close(lo.l1_loss(render, gt).item(), 0.039912, rtol=2e-4) and close(lo.ssim(render, gt).item(), 0.640930, rtol=2e-4)
