"""Host-side mirror of the reference interface: module tree, state_dict names, init RNG order, no silent fallback."""
import os

import numpy as np
import pytest
import torch

import models
from oracle import dip_oracle as O
from oracle import ref_harness

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build(mode="bilinear", seed=0):
    torch.manual_seed(seed)
    return models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                          upsample_mode=mode)


def test_state_dict_keys_match_reference():
    g = np.load(os.path.join(GOLD, "denoise64_bilinear_fp32.npz"))
    net = build()
    assert list(net.state_dict().keys()) == [str(k) for k in g["state_keys"]]
    assert sum(p.numel() for p in net.parameters()) == 2217831
    assert len(list(net.parameters())) == 112


def test_inpainting_config_is_accelerated_and_matches_reference_layout():
    """BASELINE config 4 (inpainting.ipynb kate: skip(32, 3, [128]*5, [128]*5, [128]*5, nearest, reflection))."""
    g = np.load(os.path.join(GOLD, "inpaint64x96_nearest_masked_skip128_fp32.npz"))
    torch.manual_seed(0)
    net = models.skip(32, 3, num_channels_down=[128] * 5, num_channels_up=[128] * 5, num_channels_skip=[128] * 5,
                      upsample_mode="nearest", need_sigmoid=True, need_bias=True, pad="reflection", act_fun="LeakyReLU")
    assert net._dip_spec is not None and net._dip_spec["skip_channels"] == 128 and not net._dip_spec["bilinear"]
    assert list(net.state_dict().keys()) == [str(k) for k in g["state_keys"]]
    assert sum(p.numel() for p in net.parameters()) == 3002627
    params = O.init_params(O.SkipConfig(skip_channels=128, upsample_mode="nearest"), seed=0)
    for a, b in zip(net.parameters(), params):
        assert a.shape == b.shape and torch.equal(a.detach(), b.detach())


def test_init_matches_oracle_order():
    net = build(seed=5)
    params = O.init_params(O.SkipConfig(), seed=5)
    for a, b in zip(net.parameters(), params):
        assert a.shape == b.shape and torch.equal(a.detach(), b.detach())


def test_no_silent_cpu_fallback():
    net = build()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 32, 64, 64))
    from utils.common_utils import optimize
    with pytest.raises(RuntimeError):
        optimize("adam", list(net.parameters()), lambda: None, 0.01, 1)


def test_unsupported_config_raises_not_falls_back():
    net = models.skip(3, 3, num_channels_down=[8, 16], num_channels_up=[8, 16], num_channels_skip=[0, 4],
                      upsample_mode="bilinear", pad="reflection", downsample_mode="max")   # in-net max pooling (models/common.py:106-107)
    assert net._dip_spec is None and "downsample_mode" in net._dip_why
    # per-scale widths alone (denoising.ipynb c8 "snail") ARE accelerated
    ok = models.skip(3, 3, num_channels_down=[8, 16], num_channels_up=[8, 16], num_channels_skip=[0, 4],
                     upsample_mode="bilinear", pad="reflection")
    assert ok._dip_spec is not None and ok._dip_spec["channels"] == [8, 16] and ok._dip_spec["skip_channels"] == [0, 4]
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 3, 32, 32))


def test_tree_forward_equals_oracle_when_opted_in():
    net = build(seed=3)
    z = torch.rand(1, 32, 64, 96) * 0.1
    models.allow_torch_execution(True)
    try:
        out = net(z).detach()
    finally:
        models.allow_torch_execution(False)
    params = O.init_params(O.SkipConfig(), seed=3)
    ref = O.skip_forward(params, z, O.SkipConfig()).detach()
    assert torch.allclose(out, ref, atol=1e-6)


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_tree_equals_live_reference():
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(11)
        rnet = ref.models.skip(32, 3, num_channels_down=[128] * 3, num_channels_up=[128] * 3,
                               num_channels_skip=[4] * 3, upsample_mode="nearest", pad="reflection")
        rsd = {k: v.clone() for k, v in rnet.state_dict().items()}
        z = torch.rand(1, 32, 32, 32)
        rout = rnet(z).detach()
    torch.manual_seed(11)
    net = models.skip(32, 3, num_channels_down=[128] * 3, num_channels_up=[128] * 3, num_channels_skip=[4] * 3,
                      upsample_mode="nearest", pad="reflection")
    sd = net.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    for k in sd:
        assert torch.equal(sd[k], rsd[k]), k
    models.allow_torch_execution(True)
    try:
        out = net(z).detach()
    finally:
        models.allow_torch_execution(False)
    assert torch.allclose(out, rout, atol=1e-6)


def test_get_noise_and_converters():
    from utils.common_utils import get_noise, np_to_torch, torch_to_np
    torch.manual_seed(1)
    a = get_noise(32, "noise", (16, 24))
    b = O.get_noise(32, (16, 24), seed=1)
    assert a.shape == (1, 32, 16, 24) and torch.equal(a, b)
    x = np.random.rand(3, 4, 5).astype(np.float32)
    assert np.array_equal(torch_to_np(np_to_torch(x)), x)


def test_downsampler_kernel_matches_reference_values():
    # 1-D taps of Lanczos-2, factor 4, phase 1/2 (probed from the reference)
    k = models.get_kernel(4, "lanczos", 0.5, 17, support=2)
    taps = k.sum(0)
    want = [-0.001065, -0.009752, -0.020384, -0.014878, 0.024594, 0.098658, 0.183115, 0.239711]
    assert k.shape == (16, 16) and np.allclose(taps[:8], want, atol=1e-6) and np.allclose(taps[8:], want[::-1], atol=1e-6)


def test_notebook_import_lines_resolve():
    """The import cells of the skip-net notebooks (inpainting.ipynb c3, restoration.ipynb c3, super-resolution.ipynb c3,
    flash-no-flash.ipynb c3) must resolve against this package."""
    from models import get_net, skip  # noqa: F401
    from models.downsampler import Downsampler  # noqa: F401
    from models.resnet import ResNet  # noqa: F401
    from models.skip import skip as skip2  # noqa: F401
    from models.unet import UNet  # noqa: F401
    from utils.denoising_utils import get_noisy_image  # noqa: F401
    from utils.inpainting_utils import get_bernoulli_mask, get_text_mask  # noqa: F401
    from utils.sr_utils import load_LR_HR_imgs_sr, tv_loss  # noqa: F401
    with pytest.raises(NotImplementedError):
        get_net(32, "texture_nets", "reflection", "bilinear")


UNET_CASES = [  # inpainting.ipynb c14:62-70 (library / UNET), get_net('UNet') (models/__init__.py:22-25), and the other modes
    dict(num_input_channels=1, num_output_channels=3, feature_scale=8, more_layers=1, concat_x=False, upsample_mode="deconv",
         pad="zero", norm_layer=torch.nn.InstanceNorm2d, need_sigmoid=True, need_bias=True),
    dict(num_input_channels=32, num_output_channels=3, feature_scale=4, more_layers=0, concat_x=False,
         upsample_mode="bilinear", pad="reflection", norm_layer=torch.nn.BatchNorm2d, need_sigmoid=True, need_bias=True),
    dict(num_input_channels=2, num_output_channels=1, feature_scale=16, more_layers=0, concat_x=True, upsample_mode="nearest",
         pad="zero", norm_layer=None, need_sigmoid=False, need_bias=False),
]


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
@pytest.mark.parametrize("kw", UNET_CASES)
def test_unet_builder_matches_reference(kw):
    """models.UNet keeps the reference's builder API (models/unet.py:32-192): same parameter names / shapes / init
    draws and the same output from stock torch ops."""
    torch.manual_seed(3)
    net = models.UNet(**kw)
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(3)
        rnet = ref.models.UNet(**kw)
        sd, rsd = net.state_dict(), rnet.state_dict()
        assert list(sd.keys()) == list(rsd.keys())
        for k in sd:
            assert torch.equal(sd[k], rsd[k]), k
        x = torch.rand(1, kw["num_input_channels"], 64, 96)
        want = rnet(x)
    got = net(x)
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-6)
    got.mean().backward()
    assert all(p.grad is not None for p in net.parameters())


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_resnet_builder_matches_reference():
    """inpainting.ipynb c14:72-77: ResNet(input_depth, 3, 8, 32, need_sigmoid=True, act_fun='LeakyReLU')."""
    args, kw = (1, 3, 8, 32), dict(need_sigmoid=True, act_fun="LeakyReLU")
    torch.manual_seed(4)
    net = models.ResNet(*args, **kw)
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(4)
        rnet = ref.models.ResNet(*args, **kw)
        sd, rsd = net.state_dict(), rnet.state_dict()
        assert list(sd.keys()) == list(rsd.keys()) and all(torch.equal(sd[k], rsd[k]) for k in sd)
        x = torch.rand(1, 1, 32, 48)
        want = rnet(x)
    got = net(x)
    assert torch.allclose(got, want, atol=1e-6)
    # get_net('UNet') builds; get_net('ResNet') fails exactly like the reference's (TODO-marked) call does
    assert isinstance(models.get_net(32, "UNet", "reflection", "bilinear"), models.UNet)
    with pytest.raises(TypeError):
        models.get_net(32, "ResNet", "reflection", "bilinear")
