"""Network-level parity at the BASELINE.json shapes (the configurations bench.py times), on the reference's own data:

  denoise512  denoising.ipynb c4-c10       F16 512x512, sigma 25, skip[128x5] cs=4 bilinear           (configs 1/2/5)
  inpaint512  inpainting.ipynb c5-c17      kate 512x512 + mask, skip=128 nearest, masked MSE           (config 4)
  sr_zebra    super-resolution.ipynb c5-10 zebra 384x576 -> 96x144, Lanczos-2 x4 in the loss           (config 3, real pair)
  sr1024      BASELINE wording 256 -> 1024 synthetic LR target, same operator                          (config 3)

One closure step, engine (through the C ABI) vs (a) the CPU oracle evaluated live on the same inputs -- every pre-BN
activation of every level, the output, the loss, every gradient -- and (b) the fixtures generated from the UNMODIFIED
reference (tests/golden/make_golden.py baseline -> tests/golden/baseline_*_fp32.npz).  The oracle itself is checked
against the same fixtures here (and in tests/test_oracle.py for the cases that fit the CPU suite).

Tiers: fp32 = exact-fp32 CUDA-core convs, tight; tf32 = tensor-core path (the benchmarked one), whose gradient error
is required to be like cuDNN-TF32's on the same graph (the reference's own GPU arithmetic, comparator only).
"""
import os

import numpy as np
import pytest
import torch

from oracle import dip_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
DATA = os.path.join(GOLD, "data")

FWD_TOL = {"fp32": 1e-4, "tf32": 2e-2}
RAW_TOL = {"fp32": 2e-4, "tf32": 3e-2}
GRAD_TOL = {"fp32": 3e-2}
LOSS_TOL = {"fp32": 2e-6, "tf32": 1e-3}


from baseline_cases import load_case, oracle_step, rel  # noqa: E402


def engine_step(c, prec):
    import dip_engine as de
    cfg, H, W = c["cfg"], c["H"], c["W"]
    plan = de.Plan(32, 3, cfg.num_scales, 128, cfg.skip_channels, cfg.upsample_mode == "bilinear", H, W,
                   precision=de.PRECISION_TF32 if prec == "tf32" else de.PRECISION_FP32)
    dparams = [p.detach().cuda().contiguous() for p in c["params"]]
    dgrads = [torch.zeros_like(p) for p in dparams]
    plan.bind(dparams, dgrads)
    out = plan.forward(c["z0"].cuda(), noise=c["noise"].cuda(), sigma=c["sigma"])
    L = de.lib()
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    target = c["target"].cuda().contiguous()
    mask = None if c["mask"] is None else c["mask"].cuda().contiguous()
    if c["down"] is None:
        dout = torch.empty_like(out)
        de.check(L.dip_loss_mse(out.data_ptr(), target.data_ptr(), None if mask is None else mask.data_ptr(), 3, H * W,
                                loss.data_ptr(), dout.data_ptr(), None))
    else:
        kern, f, pad = c["down"]
        kern = kern.cuda().contiguous()
        lr = de.lanczos_down_fwd(out, kern, f, pad)
        dlr = torch.empty_like(lr)
        de.check(L.dip_loss_mse(lr.data_ptr(), target.data_ptr(), None, 3, lr.shape[2] * lr.shape[3], loss.data_ptr(),
                                dlr.data_ptr(), None))
        dout = de.lanczos_down_bwd(dlr, kern, f, pad, H, W)
    plan.backward(dout)
    torch.cuda.synchronize()
    return plan, out, loss.item(), dgrads


def cudnn_tf32_grads(c):
    """The reference's own GPU arithmetic: the same graph on torch-CUDA with cuDNN's default TF32 convolutions."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        pc = [p.detach().cuda().requires_grad_(True) for p in c["params"]]
        z = (c["z0"] + c["noise"] * c["sigma"]).cuda()
        out = O.skip_forward(pc, z, c["cfg"])
        if c["down"] is not None:
            kern, f, pad = c["down"]
            x = torch.nn.functional.pad(out, (pad,) * 4, mode="replicate")
            w = torch.zeros(3, 3, *kern.shape, device="cuda")
            for i in range(3):
                w[i, i] = kern.cuda()
            o = torch.nn.functional.conv2d(x, w, stride=f)
        else:
            o = out
        m = None if c["mask"] is None else c["mask"].cuda()
        loss = O.mse_loss(o, c["target"].cuda(), m)
        return [x.detach().cpu() for x in torch.autograd.grad(loss, pc)], out.detach().cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("kind", ["denoise512", "inpaint512", "sr_zebra", "sr1024"])
def test_one_step_at_baseline_shape(kind, prec):
    c = oracle_step(kind)
    plan, out, loss, dgrads = engine_step(c, prec)
    cfg, g = c["cfg"], c["g"]
    # (1) every pre-BN activation, level by level
    for l in range(cfg.num_scales):
        for nm in ("raw_s", "raw_d1", "raw_d2", "raw_u", "raw_v"):
            key = "L%d.%s" % (l, nm)
            e = rel(plan.buffer(key), c["tape"][key][0].permute(1, 2, 0))
            assert e < RAW_TOL[prec], (key, e)
    # (2) output and loss: vs the oracle and vs the reference fixture
    assert (out.cpu() - c["out"]).abs().max().item() < FWD_TOL[prec]
    assert np.abs(out.cpu().numpy()[:, :, ::4, ::4] - g["out0_sub"]).max() < FWD_TOL[prec]
    assert abs(out.double().mean().item() - float(g["out0_mean"])) < FWD_TOL[prec] * 0.1
    assert abs(loss - c["loss"]) < LOSS_TOL[prec] and abs(loss - float(g["losses"][0])) < LOSS_TOL[prec]
    # (3) gradients
    names = [n for n, _ in O.param_layout(cfg)]
    gmax = max(x.norm().item() for x in c["grads"])
    if prec == "fp32":
        worst = ("", 0.0)
        for name, gd, gr in zip(names, dgrads, c["grads"]):
            dead = (name.endswith(".b") and "_bn" not in name and not name.startswith("head")) or name.endswith("cat_bn.b")
            if dead or gr.norm().item() < 1e-5 * gmax:
                # mathematically-zero gradients (a conv bias / the concat-BN shift in front of a BatchNorm): rounding noise
                # in the reference too, whatever its size there -- only "small" can be asserted
                assert gd.norm().item() < 1e-3 * gmax and gr.norm().item() < 1e-3 * gmax, name
                continue
            e = rel(gd, gr)
            worst = max(worst, (name, e), key=lambda t: t[1])
        assert worst[1] < GRAD_TOL["fp32"], worst
        gn = np.array([x.double().norm().item() for x in dgrads])
        big = g["gnorm0"] > 1e-4 * g["gnorm0"].max()
        assert np.abs(gn[big] / g["gnorm0"][big] - 1).max() < GRAD_TOL["fp32"]
        assert rel(dgrads[-2], torch.from_numpy(g["g_head_w"])) < GRAD_TOL["fp32"]
        assert rel(dgrads[-10][:4, :8], torch.from_numpy(g["g_up0_w_slice"])) < GRAD_TOL["fp32"]
        assert rel(dgrads[4 * 12 + 8][:4, :8], torch.from_numpy(g["g_d2_4_slice"])) < 2 * GRAD_TOL["fp32"]
    else:
        gc, out_c = cudnn_tf32_grads(c)
        e_out_ours = (out.cpu() - c["out"]).abs().max().item()
        e_out_cudnn = (out_c - c["out"]).abs().max().item()
        assert e_out_ours < 3.0 * e_out_cudnn + 2e-3, (e_out_ours, e_out_cudnn)
        e_ours, e_cudnn = [], []
        for name, gd, gcu, gr in zip(names, dgrads, gc, c["grads"]):
            dead = (name.endswith(".b") and "_bn" not in name and not name.startswith("head")) or name.endswith("cat_bn.b")
            if dead or gr.norm().item() < 1e-4 * gmax:
                continue
            eo, ec = rel(gd, gr), rel(gcu, gr)
            assert eo < 3.0 * ec + 0.08, (name, eo, ec)
            e_ours.append(eo)
            e_cudnn.append(ec)
        assert np.median(e_ours) < 1.5 * np.median(e_cudnn) + 0.01, (np.median(e_ours), np.median(e_cudnn))
    del plan
    torch.cuda.empty_cache()


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_running_stats_at_512_vs_reference_fixture(prec):
    """BatchNorm running_mean / running_var / num_batches_tracked of ALL 30 layers after the first forward of the F16
    denoising configuration through the module API, vs the unmodified reference (fixture keys rm1 / rv1 / nbt1)."""
    import models
    c = load_case("denoise512")
    g = c["g"]
    dtype = torch.cuda.FloatTensor
    torch.manual_seed(0)
    net = models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                         upsample_mode="bilinear").type(dtype)
    net.precision = prec
    with torch.no_grad():
        net((c["z0"] + c["noise"] * c["sigma"]).type(dtype))
    torch.cuda.synchronize()
    sd = net.state_dict()
    rm = np.concatenate([sd[k].cpu().numpy().ravel() for k in sd if k.endswith("running_mean")])
    rv = np.concatenate([sd[k].cpu().numpy().ravel() for k in sd if k.endswith("running_var")])
    nbt = np.array([float(sd[k]) for k in sd if k.endswith("num_batches_tracked")])
    assert rm.shape == g["rm1"].shape and np.all(nbt == g["nbt1"]) and np.all(nbt == 1)
    tol = 1e-5 if prec == "fp32" else 2e-3
    assert np.abs(rm - g["rm1"]).max() < tol * max(1.0, np.abs(g["rm1"]).max()), np.abs(rm - g["rm1"]).max()
    assert np.abs(rv / g["rv1"] - 1).max() < (1e-4 if prec == "fp32" else 2e-2), np.abs(rv / g["rv1"] - 1).max()
