"""Super-resolution operator: oracle and host-side module vs fixtures generated from the
unmodified reference (tests/golden/make_golden.py, `downsampler_cases.npz`, `sr64x96_*.npz`); on the GPU the CUDA
stencil kernels (dip_lanczos_down_fwd / _bwd through the C ABI) vs the same fixtures and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import dip_oracle as O
from oracle import ref_harness

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tag -> (oracle/down_kernel arguments, factor, preserve_size)
CASES = {
    "lanczos2_f4": (dict(factor=4, kernel_type="lanczos2", phase=0.5), 4, True),
    "lanczos2_f2": (dict(factor=2, kernel_type="lanczos2", phase=0.5), 2, True),
    "lanczos3_f4": (dict(factor=4, kernel_type="lanczos3", phase=0.5), 4, True),
    "lanczos2_f8": (dict(factor=8, kernel_type="lanczos2", phase=0.5), 8, True),
    "gauss12_f2": (dict(factor=2, kernel_type="gauss12", phase=0), 2, True),
    "box_f4": (dict(factor=4, kernel_type="box", phase=0.5, kernel_width=4), 4, True),
    "lanczos2_f4_nopad": (dict(factor=4, kernel_type="lanczos2", phase=0.5), 4, False),
}


def gold():
    return np.load(os.path.join(GOLD, "downsampler_cases.npz"))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_kernel_and_operator_match_reference_fixture(tag):
    g = gold()
    kw, f, preserve = CASES[tag]
    k = O.down_kernel(**kw)
    assert k.shape == g[tag + ".kernel"].shape and np.allclose(k, g[tag + ".kernel"], rtol=0, atol=1e-15)
    pad = O.down_pad(k.shape[0], f) if preserve else 0
    x = torch.from_numpy(g[tag + ".x"]).requires_grad_(True)
    y = O.downsample(x, k, f, pad)
    assert y.shape == g[tag + ".y"].shape and np.allclose(y.detach().numpy(), g[tag + ".y"], rtol=0, atol=1e-6)
    y.backward(torch.from_numpy(g[tag + ".dy"]))
    assert np.allclose(x.grad.numpy(), g[tag + ".dx"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_module_kernel_matches_reference_fixture(tag):
    import models
    g = gold()
    kw, f, preserve = CASES[tag]
    ds = models.Downsampler(n_planes=3, preserve_size=preserve, **kw)
    assert np.allclose(ds.kernel, g[tag + ".kernel"], rtol=0, atol=1e-15)
    assert ds.pad == (O.down_pad(ds.kernel.shape[0], f) if preserve else 0)
    w = ds.downsampler_.weight.detach()
    assert torch.equal(w[1, 1], torch.from_numpy(ds.kernel).float()) and float(w[0, 1].abs().max()) == 0.0
    # module tree executed by stock torch (opt-in) reproduces the reference's output
    models.allow_torch_execution(True)
    try:
        y = ds(torch.from_numpy(g[tag + ".x"]))
    finally:
        models.allow_torch_execution(False)
    assert np.allclose(y.detach().numpy(), g[tag + ".y"], rtol=0, atol=1e-6)
    with pytest.raises(RuntimeError):
        ds(torch.from_numpy(g[tag + ".x"]))   # no silent CPU fallback


def run_sr_oracle(g, dtype):
    H, W, f = int(g["H"]), int(g["W"]), int(g["factor"])
    cfg = O.SkipConfig(upsample_mode="bilinear")
    params = O.init_params(cfg, seed=0, dtype=dtype)
    z0 = O.get_noise(32, (H, W), seed=1).to(dtype)
    gen = torch.Generator().manual_seed(2)
    target = torch.rand(1, 3, H // f, W // f, generator=gen).to(dtype)
    gn = torch.Generator().manual_seed(123)
    noises = [torch.randn(z0.shape, generator=gn).to(dtype) for _ in range(int(g["iters"]))]
    k = O.down_kernel(f, "lanczos2", 0.5)
    rec = {}

    def record(i, out, loss, grads):
        if i == 0:
            rec["out0"], rec["grads0"] = out, [x.clone() for x in grads]

    losses, _ = O.run(cfg, params, z0, target, noises, float(g["sigma"]), float(g["lr"]), record=record,
                      down=(k, f, O.down_pad(k.shape[0], f)))
    return cfg, params, z0, target, noises, k, losses, rec


def test_sr_closure_oracle_matches_golden_fp64():
    g = np.load(os.path.join(GOLD, "sr64x96_fp64.npz"))
    _, _, _, _, _, _, losses, rec = run_sr_oracle(g, torch.float64)
    assert np.allclose(rec["out0"].numpy(), g["out0"], atol=1e-10)
    assert np.allclose(losses[0], g["losses"][0], rtol=1e-10)
    gn = np.array([x.double().norm().item() for x in rec["grads0"]])
    big = g["gnorm0"] > 1e-9
    assert np.allclose(gn[big], g["gnorm0"][big], rtol=1e-6)
    assert np.allclose(losses, g["losses"], rtol=5e-2)


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_oracle_operator_matches_live_reference():
    with ref_harness.reference_modules() as ref:
        ds = ref.models.downsampler.Downsampler(n_planes=3, factor=4, kernel_type="lanczos2", phase=0.5, preserve_size=True)
        x = torch.rand(1, 3, 40, 52)
        y_ref = ds(x).detach()
    k = O.down_kernel(4, "lanczos2", 0.5)
    assert torch.allclose(O.downsample(x, k, 4, O.down_pad(16, 4)), y_ref, atol=1e-6)


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_cuda_operator_matches_reference_fixture(tag):
    import dip_engine as de
    g = gold()
    kw, f, preserve = CASES[tag]
    k = torch.from_numpy(g[tag + ".kernel"]).float().cuda()
    pad = O.down_pad(k.shape[0], f) if preserve else 0
    x = torch.from_numpy(g[tag + ".x"]).cuda()
    y = de.lanczos_down_fwd(x, k, f, pad)
    assert tuple(y.shape) == g[tag + ".y"].shape
    assert np.abs(y.cpu().numpy() - g[tag + ".y"]).max() < 2e-6          # fp32, K*K-term sums of O(1) values
    dx = de.lanczos_down_bwd(torch.from_numpy(g[tag + ".dy"]).cuda(), k, f, pad, x.shape[2], x.shape[3])
    assert np.abs(dx.cpu().numpy() - g[tag + ".dx"]).max() < 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1024, 1024), (384, 576), (100, 36)])
def test_cuda_operator_adjoint_and_oracle_at_size(shape):
    """<A x, y> == <x, A^T y> (size-independent property, BASELINE config 3 sizes) + the oracle on the same input."""
    import dip_engine as de
    H, W = shape
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(1, 3, H, W, generator=gen)
    k64 = O.down_kernel(4, "lanczos2", 0.5)
    k = torch.from_numpy(k64).float().cuda()
    y = de.lanczos_down_fwd(x.cuda(), k, 4, 6)
    y_ref = O.downsample(x, k64, 4, 6)
    assert y.shape == y_ref.shape and (y.cpu() - y_ref).abs().max().item() < 2e-6
    dy = torch.randn(y.shape, generator=gen)
    dx = de.lanczos_down_bwd(dy.cuda(), k, 4, 6, H, W)
    lhs = (y.double().cpu() * dy.double()).sum().item()
    rhs = (x.double() * dx.double().cpu()).sum().item()
    assert abs(lhs - rhs) < 1e-5 * max(1.0, abs(lhs))


@pytest.mark.gpu
def test_module_autograd_on_gpu():
    import models
    g = gold()
    ds = models.Downsampler(n_planes=3, factor=4, kernel_type="lanczos2", phase=0.5, preserve_size=True).type(
        torch.cuda.FloatTensor)
    x = torch.from_numpy(g["lanczos2_f4.x"]).cuda().requires_grad_(True)
    y = ds(x)
    y.backward(torch.from_numpy(g["lanczos2_f4.dy"]).cuda())
    assert np.abs(y.detach().cpu().numpy() - g["lanczos2_f4.y"]).max() < 2e-6
    assert np.abs(x.grad.cpu().numpy() - g["lanczos2_f4.dx"]).max() < 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_sr_closure_through_modules_vs_golden(prec):
    """super-resolution.ipynb c8-c10 through the notebook-facing API: net + Downsampler + MSELoss + optimize()."""
    import models
    from utils.common_utils import get_noise, get_params, optimize
    g = np.load(os.path.join(GOLD, "sr64x96_fp32.npz"))
    H, W, f = int(g["H"]), int(g["W"]), int(g["factor"])
    dtype = torch.cuda.FloatTensor
    torch.manual_seed(0)
    net = models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                         upsample_mode="bilinear").type(dtype)
    net.precision = prec
    torch.manual_seed(1)
    z0 = get_noise(32, "noise", (H, W)).type(dtype).detach()
    gen = torch.Generator().manual_seed(2)
    target = torch.rand(1, 3, H // f, W // f, generator=gen).type(dtype)
    ds = models.Downsampler(n_planes=3, factor=f, kernel_type="lanczos2", phase=0.5, preserve_size=True).type(dtype)
    gn = torch.Generator().manual_seed(123)
    mse = torch.nn.MSELoss().type(dtype)
    losses, outs = [], []

    def closure():
        noise = torch.randn(z0.shape, generator=gn).type(dtype)
        out = net(z0 + noise * float(g["sigma"]))
        loss = mse(ds(out), target)
        loss.backward()
        losses.append(loss.item())
        outs.append(out.detach())
        return loss

    params = get_params("net", net, z0)
    optimize("adam", params, closure, float(g["lr"]), 1)
    gnorm = np.array([p.grad.double().norm().item() for p in params])
    tol_out, tol_loss, tol_g = (1e-4, 1e-5, 3e-2) if prec == "fp32" else (2e-2, 1e-3, 0.25)
    assert np.abs(outs[0].cpu().numpy() - g["out0"]).max() < tol_out
    assert abs(losses[0] - float(g["losses"][0])) < tol_loss
    big = g["gnorm0"] > 1e-4 * g["gnorm0"].max()
    dev = np.abs(gnorm[big] / g["gnorm0"][big] - 1)
    # tf32: at 64x96 the deepest BatchNorms normalise over 6 pixels and TF32 rounding moves single gradients by ~20%
    # (for cuDNN-TF32 too, tests/test_engine_gpu.py) -> median; fp32: every tensor
    assert (np.median(dev) if prec == "tf32" else dev.max()) < (0.1 if prec == "tf32" else tol_g), dev.max()


@pytest.mark.gpu
def test_sr_runner_matches_module_path():
    """dip_run_iterations with dip_plan_set_downsampler == forward + downsampler + MSE + backward + Adam step by step."""
    import dip_engine as de
    H, W, f = 64, 96, 4
    cfg = O.SkipConfig(upsample_mode="bilinear")
    params = O.init_params(cfg, seed=0)
    z0 = O.get_noise(32, (H, W), seed=1).cuda()
    gen = torch.Generator().manual_seed(2)
    target = torch.rand(1, 3, H // f, W // f, generator=gen).cuda()
    k64 = O.down_kernel(f, "lanczos2", 0.5)
    kd = torch.from_numpy(k64).float().cuda()

    def fresh():
        plan = de.Plan(32, 3, 5, 128, 4, True, H, W, precision=de.PRECISION_FP32)
        dparams = [p.detach().cuda().contiguous() for p in params]
        dgrads = [torch.zeros_like(p) for p in dparams]
        plan.bind(dparams, dgrads)
        for p, gbuf in zip(dparams, dgrads):
            p.grad = gbuf
        adam = de.FusedAdam(dparams, lr=0.01)
        adam._bind(dgrads)
        return plan, dparams, dgrads, adam

    # (a) runner
    plan, pa, _, adam = fresh()
    plan.set_downsampler(k64, f, 6)
    hist = torch.zeros(2, dtype=torch.float64, device="cuda")
    de.run_iterations(plan, adam, z0, target, None, 0.0, 7, 2, 0.01, loss_hist=hist)
    # (b) step by step through the single entry points
    plan2, pb, _, adam2 = fresh()
    losses = []
    for _ in range(2):
        out = plan2.forward(z0)
        y = de.lanczos_down_fwd(out, kd, f, 6)
        loss = torch.zeros(1, dtype=torch.float64, device="cuda")
        dy = torch.empty_like(y)
        de.check(de.lib().dip_loss_mse(y.data_ptr(), target.data_ptr(), None, 3, y.shape[2] * y.shape[3], loss.data_ptr(),
                                       dy.data_ptr(), None))
        plan2.backward(de.lanczos_down_bwd(dy, kd, f, 6, H, W))
        adam2.step()
        losses.append(loss.item())
    torch.cuda.synchronize()
    # oracle for the first loss
    out_ref = O.skip_forward(params, z0.cpu(), cfg)
    loss_ref = O.mse_loss(O.downsample(out_ref, k64, f, 6), target.cpu()).item()
    h = hist.cpu().numpy()
    assert abs(h[0] - loss_ref) < 1e-6 and abs(losses[0] - loss_ref) < 1e-6
    assert abs(h[1] - losses[1]) < 1e-9           # identical kernels, identical order -> identical trajectory
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
    # switching the option off restores the plain loss
    plan.set_downsampler(None, 1, 0)
    hist2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    full_target = torch.rand(1, 3, H, W, generator=gen).cuda()
    de.run_iterations(plan, adam, z0, full_target, None, 0.0, 7, 1, 0.01, loss_hist=hist2)
    torch.cuda.synchronize()
    assert np.isfinite(hist2.cpu().numpy()).all()
