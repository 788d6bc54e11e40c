"""Shared helpers of the BASELINE-shape parity tests (tests/test_baseline_shapes_gpu.py on the GPU, tests/test_oracle.py
on the CPU): inputs built with this repo's task utilities from the committed copies of the reference's images, one
oracle step, and the oracle-vs-reference-fixture assertions."""
import os

import numpy as np
import torch

from oracle import dip_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
DATA = os.path.join(GOLD, "data")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def load_case(kind):
    """Inputs built with THIS repo's task utilities (the mirrors of utils/*.py) from the committed copies of the
    reference's images; the fixture's checksums prove they equal what the reference built."""
    from utils import common_utils as cu
    from utils.denoising_utils import get_noisy_image
    from utils.sr_utils import load_LR_HR_imgs_sr
    g = np.load(os.path.join(GOLD, "baseline_%s_fp32.npz" % kind))
    mask = down = None
    cs, mode = 4, "bilinear"
    if kind == "denoise512":
        img_np = cu.pil_to_np(cu.crop_image(cu.get_image(os.path.join(DATA, "F16_GT.png"), -1)[0], d=32))
        np.random.seed(0)
        target = cu.np_to_torch(get_noisy_image(img_np, 25 / 255.)[1])
    elif kind == "inpaint512":
        img_pil = cu.crop_image(cu.get_image(os.path.join(DATA, "kate.png"), -1)[0], 64)
        mask_pil = cu.crop_image(cu.get_image(os.path.join(DATA, "kate_mask.png"), -1)[0], 64)
        target = cu.np_to_torch(cu.pil_to_np(img_pil))
        mask = cu.np_to_torch(cu.pil_to_np(mask_pil))
        assert abs(float(mask.double().sum()) - float(g["mask_sum"])) < 1e-3
        cs, mode = 128, "nearest"
    else:
        if kind == "sr_zebra":
            imgs = load_LR_HR_imgs_sr(os.path.join(DATA, "zebra_GT.png"), -1, 4, "CROP")
            target = cu.np_to_torch(imgs["LR_np"])
        else:
            gg = torch.Generator().manual_seed(2)
            target = torch.rand(1, 3, 256, 256, generator=gg)
        kern = O.down_kernel(4, "lanczos2", 0.5)
        down = (torch.from_numpy(kern).float(), 4, O.down_pad(kern.shape[0], 4))
    if "target_sum" in g:
        assert abs(float(target.double().sum()) - float(g["target_sum"])) < 1e-2, "inputs differ from the reference's"
    H, W = int(g["H"]), int(g["W"])
    cfg = O.SkipConfig(upsample_mode=mode, skip_channels=cs)
    params = O.init_params(cfg, seed=0)
    z0 = O.get_noise(32, (H, W), seed=1)
    noise = torch.randn(z0.shape, generator=torch.Generator().manual_seed(123))
    return dict(g=g, H=H, W=W, cfg=cfg, params=params, z0=z0, noise=noise, sigma=float(g["sigma"]), target=target.float(),
                mask=mask, down=down)


_ORACLE = {}


def oracle_step(kind):
    """One oracle step (cached across the precision parameters): tape of activations, output, loss, gradients."""
    if kind in _ORACLE:
        return _ORACLE[kind]
    _ORACLE.clear()     # one case at a time: a 1024^2 tape is ~10 GB of host memory
    c = load_case(kind)
    tape = {}
    z = c["z0"] + c["noise"] * c["sigma"]
    out = O.skip_forward(c["params"], z, c["cfg"], tape=tape)
    o = out if c["down"] is None else O.downsample(out, *c["down"])
    loss = O.mse_loss(o, c["target"], c["mask"])
    grads = torch.autograd.grad(loss, c["params"])
    c.update(tape={k: v.detach() for k, v in tape.items() if "raw" in k}, out=out.detach(), loss=loss.item(), grads=grads)
    # the oracle against the reference-generated fixture at this shape
    g = c["g"]
    assert abs(c["loss"] - float(g["losses"][0])) < 2e-6 * max(1.0, abs(float(g["losses"][0]))) + 1e-7
    assert np.abs(c["out"].numpy()[:, :, ::4, ::4] - g["out0_sub"]).max() < 2e-5
    gn = np.array([x.double().norm().item() for x in grads])
    big = g["gnorm0"] > 1e-4 * g["gnorm0"].max()
    assert np.abs(gn[big] / g["gnorm0"][big] - 1).max() < 3e-2   # LeakyReLU-flip floor between two fp32 CPU runs
    _ORACLE[kind] = c
    return c
