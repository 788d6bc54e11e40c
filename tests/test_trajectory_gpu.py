"""P3/P4 tiers: a 400-iteration denoising trajectory on the GPU engine vs the reference's graph on
torch-CPU (tests/golden/trajectory256.npz, two CPU runs that differ only in thread count = the reference's own spread).

The DIP trajectory is chaotic under floating-point reordering (the reference does not reproduce itself), so a pointwise
1e-3 dB criterion is not meaningful; the engine must (a) match the first iterations closely and (b) stay inside an
envelope a few times wider than the reference's own run-to-run spread, with the same noise stream."""
import os

import numpy as np
import pytest
import torch

from oracle import dip_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectory256.npz")


def problem(H, W):
    g = torch.Generator().manual_seed(2024)
    clean = torch.rand(1, 3, H // 16, W // 16, generator=g)
    clean = torch.nn.functional.interpolate(clean, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    noisy = (clean + torch.randn(clean.shape, generator=g) * (25. / 255.)).clamp(0, 1)
    return clean, noisy


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_trajectory_inside_reference_envelope(prec):
    import models
    from utils.common_utils import get_params, optimize
    g = np.load(GOLD)
    H, W, iters = int(g["H"]), int(g["W"]), int(g["iters"])
    dtype = torch.cuda.FloatTensor
    torch.manual_seed(0)
    net = models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                         upsample_mode="bilinear").type(dtype)
    net.precision = prec
    z0 = O.get_noise(32, (H, W), seed=1).type(dtype)
    clean, noisy = problem(H, W)
    noisy_d = noisy.type(dtype)
    gn = torch.Generator().manual_seed(123)
    mse = torch.nn.MSELoss().type(dtype)
    losses, psnrs = [], []

    def closure():
        z = z0 + torch.randn(z0.shape, generator=gn).cuda() * (1. / 30)     # same noise stream as the fixture
        out = net(z)
        loss = mse(out, noisy_d)
        loss.backward()
        losses.append(loss.item())
        psnrs.append(O.psnr(clean.numpy()[0], out.detach().cpu().numpy()[0]))
        return loss

    optimize("adam", get_params("net", net, z0), closure, 0.01, iters)
    la, lb, pa, pb = g["loss_a"], g["loss_b"], g["psnr_a"], g["psnr_b"]
    # (a) early iterations: same state in -> same loss
    assert abs(losses[0] - la[0]) < (1e-5 if prec == "fp32" else 2e-3)
    assert abs(losses[1] - la[1]) < 5e-3
    # (b) late: envelope = reference spread (8 vs 4 threads) widened x3, floor 0.75 dB
    for i in (100, 200, 300, iters - 1):
        lo, hi = min(pa[i], pb[i]), max(pa[i], pb[i])
        tol = max(3.0 * (hi - lo), 0.75)
        assert lo - tol <= psnrs[i] <= hi + tol, (i, psnrs[i], pa[i], pb[i])
    # the optimisation actually denoises: PSNR to the clean image ends well above the noisy input's
    noisy_psnr = O.psnr(clean.numpy()[0], noisy.numpy()[0])
    assert max(psnrs[-50:]) > noisy_psnr
    print("final PSNR engine %.3f | reference runs %.3f / %.3f" % (psnrs[-1], pa[-1], pb[-1]))
