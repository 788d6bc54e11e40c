"""Skip-net variants of the other notebooks that use the 128-wide network: flash-no-flash.ipynb c8
(3-channel image as the network input, per-scale upsampling modes) and restoration.ipynb c7 barbara (n_channels = 1,
masked loss), and inpainting.ipynb c14:1-16 "vase" (num_channels_skip = 0: no skip branch / no Concat, meshgrid input of depth
2, nearest upsampling, masked loss).  Fixtures from the unmodified reference: tests/golden/make_golden.py `variants` / `vase`."""
import os

import numpy as np
import pytest
import torch

from oracle import dip_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["flash64x96_in3_mixed", "restore64_out1_masked", "vase64x96_in2_skip0_masked", "snail64x96_in3_w8to128",
         "restorekate64x96_avg_w16to128"]
# snail: denoising.ipynb c8:13-23 -- per-scale widths [8, 16, 32, 64, 128], skips [0, 0, 0, 4, 4], 3-channel noise input
# restorekate: restoration.ipynb c7:28-36 -- widths [16, 32, 64, 128, 128], no skips, downsample_mode='avg', masked loss


def dmode(g):
    return str(g["downsample_mode"]) if "downsample_mode" in g else "stride"


def skip_ch(g):
    return int(g["skip_ch"]) if "skip_ch" in g else 4


def widths(g):
    """(num_channels_down = num_channels_up, num_channels_skip) of the fixture, per scale"""
    if "chans" in g:
        return [int(x) for x in g["chans"]], [int(x) for x in g["skips"]]
    return [128] * 5, [skip_ch(g)] * 5


def oracle_cfg(g, modes):
    chans, skips = widths(g)
    if set(chans) == {128} and len(set(skips)) == 1:
        cfg = O.SkipConfig(in_channels=int(g["in_depth"]), out_channels=int(g["out_ch"]), upsample_mode=modes, skip_channels=skips[0])
    else:
        cfg = O.SkipConfig(in_channels=int(g["in_depth"]), out_channels=int(g["out_ch"]), upsample_mode=modes, channels=chans,
                           skip_channels=skips)
    cfg.downsample_mode = dmode(g)
    return cfg


def setup(g, dtype):
    H, W = int(g["H"]), int(g["W"])
    modes = [str(m) for m in g["modes"]]
    cfg = oracle_cfg(g, modes)
    gen = torch.Generator().manual_seed(2)
    z0 = torch.rand(1, cfg.in_channels, H, W, generator=gen).to(dtype)
    if "z0" in g:   # the vase fixture's meshgrid input (the generator draws above stay, so target / mask match the fixture's)
        z0 = torch.from_numpy(g["z0"]).to(dtype)
    target = torch.rand(1, cfg.out_channels, H, W, generator=gen).to(dtype)
    mask = (torch.rand(1, 1, H, W, generator=gen) > 0.5).to(dtype) if bool(g["masked"]) else None
    gn = torch.Generator().manual_seed(123)
    noises = [torch.randn(z0.shape, generator=gn).to(dtype) for _ in range(int(g["iters"]))]
    return cfg, z0, target, mask, noises


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden_fp64(case):
    g = np.load(os.path.join(GOLD, case + "_fp64.npz"))
    cfg, z0, target, mask, noises = setup(g, torch.float64)
    params = O.init_params(cfg, seed=0, dtype=torch.float64)
    rec = {}

    def record(i, out, loss, grads):
        if i == 0:
            rec["out0"], rec["grads0"] = out, [x.clone() for x in grads]

    losses, _ = O.run(cfg, params, z0, target, noises, float(g["sigma"]), float(g["lr"]), mask=mask, record=record)
    assert np.allclose(rec["out0"].numpy(), g["out0"], atol=1e-10)
    assert np.allclose(losses[0], g["losses"][0], rtol=1e-10)
    gn = np.array([x.double().norm().item() for x in rec["grads0"]])
    big = g["gnorm0"] > 1e-9
    assert np.allclose(gn[big], g["gnorm0"][big], rtol=1e-6)
    assert np.allclose(rec["grads0"][0].numpy(), g["g_skip0_w"], rtol=1e-6, atol=1e-12)
    assert np.allclose(rec["grads0"][4 if skip_ch(g) else 0].numpy(), g["g_d1_0_w"], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("case", CASES)
def test_module_tree_matches_reference_and_is_accelerated(case):
    import models
    g = np.load(os.path.join(GOLD, case + "_fp32.npz"))
    modes = [str(m) for m in g["modes"]]
    torch.manual_seed(0)
    chans, skips = widths(g)
    net = models.skip(int(g["in_depth"]), int(g["out_ch"]), num_channels_down=chans, num_channels_up=chans,
                      num_channels_skip=skips, upsample_mode=modes, need_sigmoid=True, need_bias=True, pad="reflection",
                      downsample_mode=dmode(g))
    assert list(net.state_dict().keys()) == [str(k) for k in g["state_keys"]]
    spec = net._dip_spec
    assert spec is not None and spec["in_channels"] == int(g["in_depth"]) and spec["out_channels"] == int(g["out_ch"])
    if len(set(modes)) > 1:
        assert spec["bilinear"] == [m == "bilinear" for m in modes]
    cfg = oracle_cfg(g, modes)
    assert spec["skip_channels"] == (skip_ch(g) if set(chans) == {128} else skips)
    for a, b in zip(net.parameters(), O.init_params(cfg, seed=0)):
        assert a.shape == b.shape and torch.equal(a.detach(), b.detach())


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("case", CASES)
def test_engine_matches_reference_golden(case, prec):
    """One closure step through the notebook-facing API vs the reference's numbers (fp32 fixture)."""
    import models
    from utils.common_utils import get_params, optimize
    g = np.load(os.path.join(GOLD, case + "_fp32.npz"))
    cfg, z0, target, mask, noises = setup(g, torch.float32)
    dtype = torch.cuda.FloatTensor
    torch.manual_seed(0)
    chans, skips = widths(g)
    net = models.skip(cfg.in_channels, cfg.out_channels, num_channels_down=chans, num_channels_up=chans,
                      num_channels_skip=skips, upsample_mode=list(cfg.upsample_mode), need_sigmoid=True,
                      need_bias=True, pad="reflection", downsample_mode=dmode(g)).type(dtype)
    net.precision = prec
    z0d, tgt = z0.type(dtype), target.type(dtype)
    md = mask.type(dtype) if mask is not None else None
    mse = torch.nn.MSELoss().type(dtype)
    it = iter(noises)
    losses, outs = [], []

    def closure():
        out = net(z0d + next(it).type(dtype) * float(g["sigma"]))
        loss = mse(out * md, tgt * md) if md is not None else mse(out, tgt)
        loss.backward()
        losses.append(loss.item())
        outs.append(out.detach())
        return loss

    params = get_params("net", net, z0d)
    optimize("adam", params, closure, float(g["lr"]), 1)
    tol_out, tol_loss, tol_g = (1e-4, 1e-5, 3e-2) if prec == "fp32" else (2e-2, 1e-3, 0.25)
    assert np.abs(outs[0].cpu().numpy() - g["out0"]).max() < tol_out
    assert abs(losses[0] - float(g["losses"][0])) < tol_loss
    gnorm = np.array([p.grad.double().norm().item() for p in params])
    big = g["gnorm0"] > 1e-4 * g["gnorm0"].max()
    dev = np.abs(gnorm[big] / g["gnorm0"][big] - 1)
    assert (np.median(dev) if prec == "tf32" else dev.max()) < (0.1 if prec == "tf32" else tol_g), dev.max()

    def rel(a, b):
        a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double()
        return ((a - b).norm() / (b.norm() + 1e-30)).item()
    # the two level-0 convs that read the (zero-padded) input: 1x1 skip conv (CUDA-core path) and 3x3 stride-2 conv
    # (tf32 tier, 64 x 96: the deepest BatchNorms normalise over 2 x 3 pixels, so TF32 rounding moves the first layer's gradient by
    # tens of percent -- for cuDNN-TF32 as well, tests/test_engine_gpu.py; without skip connections every path to the first layer
    # crosses all five levels, hence the wider bound for the vase configuration)
    tol1 = 3e-2 if prec == "fp32" else (0.3 if skips[0] else 0.6)
    assert rel(params[0].grad, g["g_skip0_w"]) < tol1
    assert rel(params[4 if skips[0] else 0].grad, g["g_d1_0_w"]) < tol1
    optimize("adam", params, closure, float(g["lr"]), 2)
    assert np.isfinite(losses).all() and abs(losses[1] - float(g["losses"][1])) < 2e-2


@pytest.mark.gpu
def test_runner_with_odd_input_depth():
    """dip_run_iterations on the flash-no-flash configuration (device noise on a 3-channel input): loss decreases."""
    import dip_engine as de
    H, W = 64, 96
    cfg = O.SkipConfig(in_channels=3, out_channels=3, upsample_mode=["nearest", "nearest", "bilinear", "bilinear", "bilinear"])
    params = O.init_params(cfg, seed=0)
    plan = de.Plan(3, 3, 5, 128, 4, [m == "bilinear" for m in cfg.upsample_mode], H, W)
    dparams = [p.detach().cuda().contiguous() for p in params]
    dgrads = [torch.zeros_like(p) for p in dparams]
    plan.bind(dparams, dgrads)
    for p, gb in zip(dparams, dgrads):
        p.grad = gb
    adam = de.FusedAdam(dparams, lr=0.01)
    adam._bind(dgrads)
    gen = torch.Generator().manual_seed(4)
    z0 = torch.rand(1, 3, H, W, generator=gen).cuda()
    target = torch.rand(1, 3, H, W, generator=gen).cuda()
    hist = torch.zeros(40, dtype=torch.float64, device="cuda")
    de.run_iterations(plan, adam, z0, target, None, 0.03, 7, 40, 0.01, loss_hist=hist)
    torch.cuda.synchronize()
    h = hist.cpu().numpy()
    out_ref = O.skip_forward(params, z0.cpu(), cfg)          # first loss without noise is close to the noisy one
    assert np.all(np.isfinite(h)) and h[-5:].mean() < h[:5].mean()
    assert abs(h[0] - O.mse_loss(out_ref, target.cpu()).item()) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("cfgargs", [dict(), dict(in_channels=3), dict(skip_channels=128, upsample_mode="nearest"),
                                     dict(skip_channels=0, in_channels=2, upsample_mode="nearest")])
def test_input_gradient_and_no_sigmoid_vs_oracle(cfgargs):
    """OPT_OVER = 'net,input' (utils/common_utils.py:47-49: the input tensor is optimised too) and need_sigmoid=False
    (models/skip.py:97): dL/d(net_input), the output and the weight gradients vs autograd of the oracle, exact-fp32 mode."""
    import models
    H, W = 64, 96
    kw = dict(in_channels=32, skip_channels=4, upsample_mode="bilinear")
    kw.update(cfgargs)
    cfg = O.SkipConfig(need_sigmoid=False, **kw)
    params = O.init_params(cfg, seed=0)
    gen = torch.Generator().manual_seed(2)
    z0 = (torch.rand(1, cfg.in_channels, H, W, generator=gen) * 0.1).requires_grad_(True)
    target = torch.rand(1, 3, H, W, generator=gen)
    out_ref = O.skip_forward(params, z0, cfg)
    loss_ref = O.mse_loss(out_ref, target)
    grads_ref = torch.autograd.grad(loss_ref, [z0] + params)
    dz_ref, gw_ref = grads_ref[0], grads_ref[1:]

    torch.manual_seed(0)
    net = models.skip(cfg.in_channels, 3, num_channels_down=[128] * 5, num_channels_up=[128] * 5,
                      num_channels_skip=[cfg.skip_channels] * 5, upsample_mode=cfg.upsample_mode, need_sigmoid=False,
                      need_bias=True, pad="reflection").type(torch.cuda.FloatTensor)
    net.precision = "fp32"
    zd = z0.detach().cuda().requires_grad_(True)
    out = net(zd)
    loss = torch.nn.functional.mse_loss(out, target.cuda())
    loss.backward()
    assert (out.detach().cpu() - out_ref.detach()).abs().max().item() < 5e-4
    assert abs(loss.item() - loss_ref.item()) < 1e-5

    def rel(a, b):
        a, b = a.double().cpu(), b.double()
        return ((a - b).norm() / (b.norm() + 1e-30)).item()
    assert zd.grad is not None and zd.grad.shape == z0.shape
    assert rel(zd.grad, dz_ref) < 3e-2, rel(zd.grad, dz_ref)
    gmax = max(g.norm().item() for g in gw_ref)
    for p, g in zip(net.parameters(), gw_ref):
        if g.norm().item() > 1e-4 * gmax:
            assert rel(p.grad, g) < 3e-2
    # the plain path (input does not require grad) still works on the same network object, tf32 default mode too
    net.precision = "tf32"
    out2 = net(z0.detach().cuda())
    assert (out2.detach().cpu() - out_ref.detach()).abs().max().item() < 5e-2
