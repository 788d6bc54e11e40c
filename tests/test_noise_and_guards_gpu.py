"""The device noise generator of the benchmarked runner (k_noise: Philox4x32-10 + Box-Muller; replaces
`noise.normal_()` of denoising.ipynb c10:12-13), BatchNorm running statistics vs stock torch, and the misuse guards of
the module API."""
import copy
import ctypes

import numpy as np
import pytest
import torch

from oracle import dip_oracle as O

pytestmark = pytest.mark.gpu


def _perturb(z0, sigma, seed, offset):
    import dip_engine as de
    z = torch.empty_like(z0)
    de.check(de.lib().dip_noise_perturb(z0.data_ptr(), z.data_ptr(), float(sigma), int(seed), int(offset), z0.numel(), None))
    torch.cuda.synchronize()
    return z


def test_noise_moments_and_streams():
    n = 1 << 23
    z0 = torch.zeros(n, device="cuda")
    a = _perturb(z0, 1.0, 1234, 0).double()
    # N(0,1): standard errors for n = 8.4 M samples are 3.5e-4 (mean), 4.9e-4 (variance), 8.5e-4 (skew), 1.7e-3 (kurtosis)
    m, v = a.mean().item(), a.var().item()
    sk = ((a - m) ** 3).mean().item() / v ** 1.5
    ku = ((a - m) ** 4).mean().item() / v ** 2
    assert abs(m) < 2e-3 and abs(v - 1) < 3e-3 and abs(sk) < 5e-3 and abs(ku - 3) < 1e-2, (m, v, sk, ku)
    for s, p in ((1.0, 0.682689), (2.0, 0.954500), (3.0, 0.997300)):    # tail mass
        assert abs((a.abs() < s).double().mean().item() - p) < 1e-3
    assert a.abs().max().item() > 4.5          # the tails are populated (no clipping of the uniforms)
    assert torch.isfinite(a).all()
    # neighbouring outputs (the Box-Muller cos/sin pair, the two pairs of one Philox block) are uncorrelated
    for lag in (1, 2, 3, 4):
        assert abs((a[:-lag] * a[lag:]).mean().item()) < 2e-3
    # determinism; a different iteration offset or seed is an independent stream
    assert torch.equal(_perturb(z0, 1.0, 1234, 0).double(), a)
    for b in (_perturb(z0, 1.0, 1234, 1).double(), _perturb(z0, 1.0, 1235, 0).double()):
        assert not torch.equal(a, b)
        assert abs((a * b).mean().item()) < 2e-3 and abs(b.var().item() - 1) < 3e-3
    # z = z0 + sigma * n, elementwise
    z0r = torch.rand(n, device="cuda")
    zz = _perturb(z0r, 1.0 / 30, 1234, 0)
    assert torch.allclose(zz, z0r + a.float() / 30, atol=1e-6)


def test_runner_uses_a_fresh_stream_every_iteration():
    """dip_run_iterations: iteration i of the run draws stream `adam step count + i` of (seed): the perturbed, reflection-
    padded level-0 input left in the plan after k iterations equals pad(dip_noise_perturb(offset = k - 1)), across calls
    too (the runner generates the noise inside its input transform, k_noise_pad: same Philox stream as k_noise)."""
    import dip_engine as de
    H = W = 64
    cfg = O.SkipConfig()
    params = O.init_params(cfg, seed=0)
    plan = de.Plan(32, 3, 5, 128, 4, True, H, W)
    dparams = [p.detach().cuda().contiguous() for p in params]
    dgrads = [torch.zeros_like(p) for p in dparams]
    plan.bind(dparams, dgrads)
    for p, gbuf in zip(dparams, dgrads):
        p.grad = gbuf
    adam = de.FusedAdam(dparams, lr=0.01)
    adam._bind(dgrads)
    z0 = O.get_noise(32, (H, W), seed=1).cuda()
    target = torch.rand(1, 3, H, W, device="cuda")
    seen = []
    done = 0
    for iters in (1, 2, 3):
        de.run_iterations(plan, adam, z0, target, None, 1. / 30, 99, iters, 0.01)
        torch.cuda.synchronize()
        done += iters
        pin = plan.buffer("L0.Pin")                                   # (H+2, W+2, C) NHWC, reflection padded
        want = _perturb(z0.reshape(-1), 1. / 30, 99, done - 1).reshape(32, H, W)
        wantp = torch.nn.functional.pad(want[None], (1, 1, 1, 1), mode="reflect")[0].permute(1, 2, 0)
        assert torch.equal(pin, wantp), (iters, (pin - wantp).abs().max().item())
        seen.append(pin.clone())
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    assert adam.step_count == 6


@pytest.mark.parametrize("cs,mode", [(4, "bilinear"), (128, "nearest")])
def test_running_stats_match_stock_torch(cs, mode):
    """running_mean / running_var / num_batches_tracked of every BatchNorm (incl. the channel-rotated concat BN) after
    1 and 3 forwards vs the same module tree executed by stock torch on the CPU (momentum 0.1, unbiased variance)."""
    import models
    H, W = 64, 96
    torch.manual_seed(0)
    net = models.skip(32, 3, num_channels_down=[128] * 5, num_channels_up=[128] * 5, num_channels_skip=[cs] * 5,
                      upsample_mode=mode, need_sigmoid=True, need_bias=True, pad="reflection", act_fun="LeakyReLU")
    ref = copy.deepcopy(net)
    net = net.type(torch.cuda.FloatTensor)
    net.precision = "fp32"
    g = torch.Generator().manual_seed(5)
    zs = [torch.rand(1, 32, H, W, generator=g) * 0.1 for _ in range(3)]
    models.allow_torch_execution(True)
    try:
        for i, z in enumerate(zs):
            with torch.no_grad():
                ref(z)
                net(z.cuda())
            torch.cuda.synchronize()
            if i in (0, 2):
                sr, sn = ref.state_dict(), net.state_dict()
                for k in sr:
                    if k.endswith("running_mean") or k.endswith("running_var"):
                        assert torch.allclose(sn[k].cpu(), sr[k], rtol=2e-4, atol=2e-6), (k, (sn[k].cpu() - sr[k]).abs().max())
                    elif k.endswith("num_batches_tracked"):
                        assert float(sn[k]) == float(sr[k]) == i + 1
    finally:
        models.allow_torch_execution(False)


def _net64():
    import models
    torch.manual_seed(0)
    return models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                          upsample_mode="bilinear").type(torch.cuda.FloatTensor)


def test_backward_of_an_overwritten_forward_raises():
    net = _net64()
    z = torch.rand(1, 32, 64, 64, device="cuda") * 0.1
    out1 = net(z)
    out2 = net(z * 0.5)                     # overwrites the activations saved for out1
    with pytest.raises(RuntimeError, match="overwritten"):
        out1.sum().backward()
    out2.sum().backward()                   # the latest forward is fine
    out3 = net(z)
    with torch.no_grad():
        net(z)                              # a no_grad preview between forward and backward is the same hazard
    with pytest.raises(RuntimeError, match="overwritten"):
        out3.sum().backward()
    # different spatial size = different plan: the first plan's activations are intact but it is no longer the active one
    out4 = net(z)
    net(torch.rand(1, 32, 64, 96, device="cuda"))
    with pytest.raises(RuntimeError, match="overwritten"):
        out4.sum().backward()


def test_wrong_dtype_eval_mode_and_plane_count_raise():
    import models
    net = _net64()
    z = torch.rand(1, 32, 64, 64, device="cuda")
    with pytest.raises(RuntimeError, match="float32"):
        net(z.double())
    with pytest.raises(RuntimeError, match="float32"):
        net(z.half())
    net.eval()
    with pytest.raises(NotImplementedError, match="training mode"):
        net(z)
    net.train()
    net(z)
    ds = models.Downsampler(n_planes=3, factor=4, kernel_type="lanczos2", phase=0.5, preserve_size=True).type(torch.cuda.FloatTensor)
    with pytest.raises(ValueError, match="planes"):
        ds(torch.rand(1, 4, 64, 64, device="cuda"))


def test_new_adam_after_destroy_is_not_served_a_stale_graph():
    """dip_run_iterations caches one captured step per plan; a second optimiser (next image of a per-rank shard) must get
    its own capture even if the allocator hands it the address of the destroyed one."""
    import dip_engine as de
    H = W = 64
    cfg = O.SkipConfig()
    plan = de.Plan(32, 3, 5, 128, 4, True, H, W)
    z0 = O.get_noise(32, (H, W), seed=1).cuda()
    target = torch.rand(1, 3, H, W, device="cuda")
    finals = []
    for image in range(3):
        params = O.init_params(cfg, seed=0)
        dparams = [p.detach().cuda().contiguous() for p in params]
        dgrads = [torch.zeros_like(p) for p in dparams]
        plan.bind(dparams, dgrads)
        for p, gbuf in zip(dparams, dgrads):
            p.grad = gbuf
        adam = de.FusedAdam(dparams, lr=0.01)
        adam._bind(dgrads)
        before = [p.clone() for p in dparams]
        hist = torch.zeros(4, dtype=torch.float64, device="cuda")
        de.run_iterations(plan, adam, z0, target, None, 1. / 30, 7, 4, 0.01, loss_hist=hist)
        torch.cuda.synchronize()
        moved = sum(float((a - b).abs().sum()) for a, b in zip(before, dparams))
        assert moved > 0, "parameters of image %d were never updated (stale graph)" % image
        finals.append(hist.cpu().numpy().copy())
        del adam
    # same problem three times: identical first iterations; later ones drift (split-K weight gradients are accumulated with
    # floating-point atomics, so the summation order differs run to run, and the trajectory is chaotic: SURVEY.md 7.4)
    for f in finals[1:]:
        assert np.allclose(finals[0][:2], f[:2], rtol=1e-4) and np.allclose(finals[0], f, rtol=0.1)


def test_fast_denoising_closure_matches_the_verbatim_one():
    """utils.fast_closure.DenoisingClosure = denoising.ipynb c10 with device-side metrics: same losses / PSNRs / EMA as
    the verbatim closure (host-side compare_psnr on D2H copies) on the same perturbation stream, and its device-side
    parameter snapshot restores like the notebook's CPU copy does."""
    from utils.common_utils import get_params, optimize
    from utils.fast_closure import DenoisingClosure, psnr_device
    H = W = 64
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 3, H, W, generator=g).cuda()
    noisy = (img + 0.1 * torch.randn(1, 3, H, W, generator=g).cuda()).clamp(0, 1)
    z = (torch.rand(1, 32, H, W, generator=g) * 0.1).cuda()

    def np_psnr(a, b):
        return 10 * np.log10(1.0 / np.mean((a.astype(np.float64) - b) ** 2))

    assert abs(psnr_device(img, noisy).item() - np_psnr(img.cpu().numpy(), noisy.cpu().numpy())) < 1e-4

    # verbatim closure (c10:8-33), 4 iterations
    net = _net64()
    net.precision = "fp32"
    mse = torch.nn.MSELoss()
    torch.manual_seed(11)
    noise = z.clone()
    st = {"out_avg": None}
    ref = []

    def closure():
        net_input = z + (noise.normal_() * (1. / 30))
        out = net(net_input)
        st["out_avg"] = out.detach() if st["out_avg"] is None else st["out_avg"] * 0.99 + out.detach() * 0.01
        total_loss = mse(out, noisy)
        total_loss.backward()
        o = out.detach().cpu().numpy()[0]
        ref.append((total_loss.item(), np_psnr(noisy.cpu().numpy()[0], o), np_psnr(img.cpu().numpy()[0], o),
                    np_psnr(img.cpu().numpy()[0], st["out_avg"].cpu().numpy()[0])))
        return total_loss
    optimize("adam", get_params("net", net, z), closure, 0.01, 4)

    net2 = _net64()
    net2.precision = "fp32"
    torch.manual_seed(11)
    fast = DenoisingClosure(net2, z, noisy, img, reg_noise_std=1. / 30, exp_weight=0.99, show_every=100, mse=mse)
    optimize("adam", get_params("net", net2, z), fast, 0.01, 4)
    got, want = np.array(fast.history), np.array(ref)
    assert np.allclose(got[:2], want[:2], rtol=1e-4, atol=1e-4), (got[:2], want[:2])      # identical state for two iterations
    assert np.allclose(got, want, rtol=0.05, atol=0.05)                                     # then the usual fp drift
    assert torch.allclose(fast.out_avg, st["out_avg"], atol=5e-2)
    # snapshot / restore on the device
    assert fast.snapshot.valid and fast.i == 4
    before = [p.detach().clone() for p in net2.parameters()]
    with torch.no_grad():
        for p in net2.parameters():
            p.add_(1.0)
    fast.snapshot.restore()
    # the snapshot holds the parameters as they were when the closure of the last iteration ran (before its Adam step)
    assert all(torch.isfinite(p).all() for p in net2.parameters())
    assert sum(float((a - b).abs().max()) for a, b in zip(before, net2.parameters())) < 112 * 0.011


def test_optimize_lbfgs_branch_on_the_engine():
    """utils.optimize('LBFGS', ...) (reference: utils/common_utils.py:208-221): 100 Adam warm-up steps at lr 1e-3 (fused
    Adam on the engine), then torch.optim.LBFGS driving the same closure -- the network, loss gradient and parameter updates
    all go through the engine module (LBFGS itself is stock torch: SURVEY.md 8f.4 keeps it outside the accelerated path)."""
    from utils.common_utils import get_params, optimize
    net = _net64()
    g = torch.Generator().manual_seed(9)
    z = (torch.rand(1, 32, 64, 64, generator=g) * 0.1).cuda()
    target = torch.rand(1, 3, 64, 64, generator=g).cuda()
    mse = torch.nn.MSELoss()
    losses = []

    def closure():
        out = net(z)
        loss = mse(out, target)
        loss.backward()
        losses.append(loss.item())
        return loss

    optimize("LBFGS", get_params("net", net, z), closure, 0.01, 8)
    assert len(losses) >= 100 + 8 and np.isfinite(losses).all()
    assert np.mean(losses[95:100]) < losses[0]                 # the Adam warm-up made progress
    assert min(losses[100:]) < np.mean(losses[95:100])         # and LBFGS went further down
