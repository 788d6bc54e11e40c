"""Pins oracle/dip_oracle.py against the reference: committed golden fixtures (generated from the unmodified
reference by tests/golden/make_golden.py) and, when /root/reference is present, the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import dip_oracle as O
from oracle import ref_harness

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def run_oracle(g, dtype, iters=None):
    H, W = int(g["H"]), int(g["W"])
    cfg = O.SkipConfig(upsample_mode=str(g["mode"]), skip_channels=int(g["skip_n11"]) if "skip_n11" in g else 4)
    params = O.init_params(cfg, seed=0, dtype=dtype)
    z0 = O.get_noise(32, (H, W), seed=1).to(dtype)
    gen = torch.Generator().manual_seed(2)
    target = torch.rand(1, 3, H, W, generator=gen).to(dtype)
    mask = (torch.rand(1, 1, H, W, generator=gen) > 0.3).to(dtype) if bool(g["masked"]) else None
    gn = torch.Generator().manual_seed(123)
    n = int(g["iters"]) if iters is None else iters
    noises = [torch.randn(z0.shape, generator=gn).to(dtype) for _ in range(n)]
    rec = {}

    def record(i, out, loss, grads):
        if i == 0:
            rec["out0"] = out
            rec["grads0"] = [x.clone() for x in grads]

    losses, _ = O.run(cfg, params, z0, target, noises, float(g["sigma"]), float(g["lr"]), mask=mask, record=record)
    return cfg, params, losses, rec


@pytest.mark.parametrize("name", ["denoise64_bilinear_fp64", "denoise96x64_nearest_masked_fp64",
                                  "inpaint64x96_nearest_masked_skip128_fp64"])
def test_oracle_matches_golden_fp64(name):
    torch.set_num_threads(8)
    g = load(name)
    cfg, params, losses, rec = run_oracle(g, torch.float64)
    # fp64: the restated graph must reproduce the reference to rounding
    assert np.allclose(rec["out0"].numpy(), g["out0"], atol=1e-10)
    assert np.allclose(losses[0], g["losses"][0], rtol=1e-10)
    gn = np.array([x.double().norm().item() for x in rec["grads0"]])
    big = g["gnorm0"] > 1e-9   # conv biases in front of a BatchNorm have mathematically-zero gradients
    assert np.allclose(gn[big], g["gnorm0"][big], rtol=1e-6)
    assert np.allclose(rec["grads0"][-2].numpy(), g["g_head_w"], rtol=1e-6, atol=1e-12)
    assert np.allclose(rec["grads0"][-10][:4, :8].numpy(), g["g_up0_w_slice"], rtol=1e-5, atol=1e-12)
    # later iterations: chaotic amplification of rounding noise (zero-gradient biases under Adam) -> loose
    assert np.allclose(losses, g["losses"], rtol=5e-2)


def test_oracle_matches_golden_fp32():
    torch.set_num_threads(8)
    g = load("denoise64_bilinear_fp32")
    cfg, params, losses, rec = run_oracle(g, torch.float32, iters=2)
    assert np.allclose(rec["out0"].numpy(), g["out0"], atol=2e-6)
    assert abs(losses[0] - g["losses"][0]) < 1e-6
    assert abs(losses[1] - g["losses"][1]) < 2e-3


def test_inpaint_param_layout_matches_reference():
    g = load("inpaint64x96_nearest_masked_skip128_fp64")
    keys = [k for k in g["state_keys"] if not ("running" in k or "num_batches" in k)]
    lay = O.param_layout(O.SkipConfig(skip_channels=128, upsample_mode="nearest"))
    assert len(lay) == len(keys) == 112
    assert sum(int(np.prod(s)) for _, s in lay) == 3002627


def test_param_layout_matches_reference_state_dict():
    g = load("denoise64_bilinear_fp32")
    keys = [k for k in g["state_keys"] if not ("running" in k or "num_batches" in k)]
    cfg = O.SkipConfig()
    lay = O.param_layout(cfg)
    assert len(lay) == len(keys) == 112
    assert sum(int(np.prod(s)) for _, s in lay) == 2217831


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_oracle_matches_live_reference():
    torch.set_num_threads(8)
    cfg = O.SkipConfig(upsample_mode="bilinear")
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(7)
        net = ref.models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                                 upsample_mode="bilinear")
        ref_params = [p.detach().clone() for p in net.parameters()]
        z = torch.rand(1, 32, 64, 96) * 0.1
        out_ref = net(z).detach()
    params = O.init_params(cfg, seed=7)
    for a, b in zip(params, ref_params):
        assert a.shape == b.shape and torch.equal(a.detach(), b)   # same init values AND same RNG order
    out = O.skip_forward(params, z, cfg).detach()
    assert torch.allclose(out, out_ref, atol=1e-6)


@pytest.mark.parametrize("kind", ["denoise512", "inpaint512", "sr_zebra"])
def test_oracle_matches_reference_fixture_at_baseline_shape(kind):
    """One step at the BASELINE.json shapes on the reference's own data (F16 512x512; kate 512x512 + mask, skip=128;
    zebra 384x576 with the Lanczos downsampler): loss, output (stride-4 grid) and gradient norms of the oracle vs the
    fixture written by the unmodified reference (tests/golden/make_golden.py baseline).  The assertions live in
    baseline_cases.oracle_step; sr1024 is checked on the GPU box only (a 1024x1024 CPU step costs a minute here)."""
    from baseline_cases import oracle_step
    c = oracle_step(kind)
    assert np.isfinite(c["loss"])


def test_bf16_operand_mode_of_the_oracle_is_a_faithful_conv_when_rounding_is_the_identity(monkeypatch):
    """The checker of the engine's bf16 mode (`with O.operand_rounding('bf16')`) swaps the wide convolutions for a custom
    autograd function (operands rounded in forward AND backward).  With the rounding replaced by the identity it must
    reproduce the plain oracle exactly -- output and every gradient -- on a per-scale-width network with skip branches;
    with real rounding it must change the wide convolutions only (the 4-output skip convs and the head stay fp32)."""
    cfg = O.SkipConfig(in_channels=3, num_scales=3, channels=[8, 16, 32], skip_channels=[0, 4, 4])
    params = O.init_params(cfg, seed=0, dtype=torch.float64)
    z = O.get_noise(3, (32, 48), seed=1).double()
    target = torch.rand(1, 3, 32, 48, generator=torch.Generator().manual_seed(2)).double()
    out0 = O.skip_forward(params, z, cfg)
    g0 = torch.autograd.grad(O.mse_loss(out0, target), params)
    monkeypatch.setattr(O, "_round_bf16", lambda t: t)
    with O.operand_rounding("bf16"):
        out1 = O.skip_forward(params, z, cfg)
        g1 = torch.autograd.grad(O.mse_loss(out1, target), params)
    assert torch.allclose(out0, out1, atol=1e-13)
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-14)
    monkeypatch.undo()
    # real rounding: a lone wide conv sees bf16 operands, a 4-output conv does not
    x = torch.randn(1, 8, 6, 6, dtype=torch.float64)
    w_wide, w_skinny = torch.randn(16, 8, 1, 1, dtype=torch.float64), torch.randn(4, 8, 1, 1, dtype=torch.float64)
    with O.operand_rounding("bf16"):
        y_wide, y_skinny = O._conv(x, w_wide, None), O._conv(x, w_skinny, None)
    assert torch.equal(y_skinny, torch.nn.functional.conv2d(x, w_skinny))
    assert torch.equal(y_wide, torch.nn.functional.conv2d(x.bfloat16().double(), w_wide.bfloat16().double()))
    assert not torch.equal(y_wide, torch.nn.functional.conv2d(x, w_wide))
    assert O._OPERAND_ROUND is None
