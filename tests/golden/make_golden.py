"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on torch-CPU.

Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin oracle/dip_oracle.py (tests/test_oracle.py) and, on the GPU box, the CUDA engine
(tests/test_engine_gpu.py).  Everything is seeded here because the reference seeds nothing.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(name, H, W, mode, iters, sigma=1. / 30, lr=0.01, masked=False, dtype=torch.float32, threads=8, skip_n11=4):
    torch.set_num_threads(threads)
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(0)
        net = ref.models.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=skip_n11, num_scales=5,
                                 upsample_mode=mode).type(dtype)
        torch.manual_seed(1)
        z0 = ref.common_utils.get_noise(32, 'noise', (H, W)).type(dtype).detach()
        g = torch.Generator().manual_seed(2)
        target = torch.rand(1, 3, H, W, generator=g).type(dtype)
        mask = None
        if masked:
            mask = (torch.rand(1, 1, H, W, generator=g) > 0.3).type(dtype)
        gn = torch.Generator().manual_seed(123)
        mse = torch.nn.MSELoss()
        params = ref.common_utils.get_params('net', net, z0)
        opt = torch.optim.Adam(params, lr=lr)
        losses, out0, gnorm0, gsum0 = [], None, None, None
        for i in range(iters):
            noise = torch.randn(z0.shape, generator=gn).type(dtype)
            opt.zero_grad()
            out = net(z0 + noise * sigma)
            loss = mse(out * mask, target * mask) if masked else mse(out, target)
            loss.backward()
            if i == 0:
                out0 = out.detach().clone()
                gnorm0 = np.array([p.grad.double().norm().item() for p in params])
                gsum0 = np.array([p.grad.double().sum().item() for p in params])
                g_head_w = params[-2].grad.detach().clone().numpy()
                g_up0_w_slice = params[-10].grad.detach()[:4, :8].clone().numpy()  # L0.up.w[:4,:8]
            losses.append(loss.item())
            opt.step()
        pnorm = np.array([p.detach().double().norm().item() for p in params])
        rm = net.state_dict()['4.running_mean'].numpy().copy()     # BN after L0.up conv
        rv = net.state_dict()['4.running_var'].numpy().copy()
        nbt = int(net.state_dict()['4.num_batches_tracked'])
        keys = list(net.state_dict().keys())
    np.savez_compressed(os.path.join(HERE, name + '.npz'), H=H, W=W, mode=mode, iters=iters, sigma=sigma, lr=lr,
                        masked=masked, losses=np.array(losses), out0=out0.numpy(), gnorm0=gnorm0, gsum0=gsum0,
                        g_head_w=g_head_w, g_up0_w_slice=g_up0_w_slice, pnorm=pnorm, rm=rm, rv=rv, nbt=nbt,
                        dtype=str(dtype), state_keys=np.array(keys), skip_n11=skip_n11)
    print(name, 'losses', losses)


def run_variant(name, H, W, in_depth, out_ch, modes, iters=3, sigma=0.03, lr=0.01, masked=False, dtype=torch.float32,
                threads=8, skip_ch=4, meshgrid=False, chans=None, skips=None, downsample_mode='stride'):
    """Skip-net variants of the other notebooks that use the 128-wide network: flash-no-flash.ipynb c8 (image as input,
    per-scale upsampling modes) and restoration.ipynb c7 barbara (n_channels=1, masked loss)."""
    torch.set_num_threads(threads)
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(0)
        chans = list(chans) if chans is not None else [128] * 5      # per-scale widths: denoising.ipynb c8:17-23 "snail"
        skips = list(skips) if skips is not None else [skip_ch] * 5
        skip_ch = skips[0]
        net = ref.models.skip(in_depth, out_ch, num_channels_down=chans, num_channels_up=chans,
                              num_channels_skip=skips, upsample_mode=modes, need_sigmoid=True, need_bias=True,
                              pad='reflection', downsample_mode=downsample_mode).type(dtype)
        g = torch.Generator().manual_seed(2)
        z0 = torch.rand(1, in_depth, H, W, generator=g).type(dtype)          # an image (or noise) as the network input
        if meshgrid:   # inpainting.ipynb c14:1-16 (vase): INPUT = 'meshgrid', input_depth = 2 (utils/common_utils.py:145-149)
            z0 = ref.common_utils.get_noise(in_depth, 'meshgrid', (H, W)).type(dtype)
        target = torch.rand(1, out_ch, H, W, generator=g).type(dtype)
        mask = (torch.rand(1, 1, H, W, generator=g) > 0.5).type(dtype) if masked else None
        gn = torch.Generator().manual_seed(123)
        mse = torch.nn.MSELoss()
        params = [p for p in net.parameters()]
        opt = torch.optim.Adam(params, lr=lr)
        losses = []
        for i in range(iters):
            noise = torch.randn(z0.shape, generator=gn).type(dtype)
            opt.zero_grad()
            out = net(z0 + noise * sigma)
            loss = mse(out * mask, target * mask) if masked else mse(out, target)
            loss.backward()
            if i == 0:
                out0 = out.detach().clone()
                gnorm0 = np.array([p.grad.double().norm().item() for p in params])
                g_first = [params[k].grad.detach().clone().numpy() for k in (0, 4 if skip_ch else 0)]   # L0 skip conv w (if any), L0 down conv w
            losses.append(loss.item())
            opt.step()
        keys = list(net.state_dict().keys())
    np.savez_compressed(os.path.join(HERE, name + '.npz'), H=H, W=W, in_depth=in_depth, out_ch=out_ch,
                        modes=np.array(modes if isinstance(modes, list) else [modes] * 5), iters=iters, sigma=sigma, lr=lr,
                        masked=masked, losses=np.array(losses), out0=out0.numpy(), gnorm0=gnorm0, g_skip0_w=g_first[0],
                        g_d1_0_w=g_first[1], dtype=str(dtype), state_keys=np.array(keys), skip_ch=skip_ch, z0=z0.numpy(),
                        chans=np.array(chans), skips=np.array(skips), downsample_mode=downsample_mode)
    print(name, 'losses', losses)


FLASH_MODES = ['nearest', 'nearest', 'bilinear', 'bilinear', 'bilinear']

DOWN_CASES = [  # (tag, ctor kwargs, H, W)
    ('lanczos2_f4', dict(factor=4, kernel_type='lanczos2', phase=0.5, preserve_size=True), 64, 96),
    ('lanczos2_f2', dict(factor=2, kernel_type='lanczos2', phase=0.5, preserve_size=True), 38, 50),
    ('lanczos3_f4', dict(factor=4, kernel_type='lanczos3', phase=0.5, preserve_size=True), 64, 64),
    ('lanczos2_f8', dict(factor=8, kernel_type='lanczos2', phase=0.5, preserve_size=True), 64, 128),
    ('gauss12_f2', dict(factor=2, kernel_type='gauss12', phase=0, preserve_size=True), 33, 47),
    ('box_f4', dict(factor=4, kernel_type='box', phase=0.5, kernel_width=4, preserve_size=True), 32, 40),
    ('lanczos2_f4_nopad', dict(factor=4, kernel_type='lanczos2', phase=0.5, preserve_size=False), 45, 70),
]


def run_downsampler_cases():
    """Forward + input gradient of the reference's Downsampler on seeded inputs (fp32, as the notebooks run it)."""
    out = {}
    with ref_harness.reference_modules() as ref:
        for tag, kw, H, W in DOWN_CASES:
            ds = ref.models.downsampler.Downsampler(n_planes=3, **kw).type(torch.FloatTensor)
            g = torch.Generator().manual_seed(11)
            x = torch.rand(1, 3, H, W, generator=g).requires_grad_(True)
            y = ds(x)
            dy = torch.randn(y.shape, generator=g)
            y.backward(dy)
            out[tag + '.kernel'] = ds.kernel
            out[tag + '.x'] = x.detach().numpy()
            out[tag + '.y'] = y.detach().numpy()
            out[tag + '.dy'] = dy.numpy()
            out[tag + '.dx'] = x.grad.numpy()
            print(tag, 'kernel', ds.kernel.shape, 'y', tuple(y.shape))
    np.savez_compressed(os.path.join(HERE, 'downsampler_cases.npz'), **out)


def run_sr_case(name, H, W, iters, dtype, factor=4, sigma=0.03, lr=0.01, threads=8):
    """super-resolution.ipynb c8-c10 on a small synthetic pair: loss = mse(downsampler(net(z)), img_LR)."""
    torch.set_num_threads(threads)
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(0)
        net = ref.models.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                                 upsample_mode='bilinear').type(dtype)
        torch.manual_seed(1)
        z0 = ref.common_utils.get_noise(32, 'noise', (H, W)).type(dtype).detach()
        g = torch.Generator().manual_seed(2)
        target = torch.rand(1, 3, H // factor, W // factor, generator=g).type(dtype)
        ds = ref.models.downsampler.Downsampler(n_planes=3, factor=factor, kernel_type='lanczos2', phase=0.5,
                                    preserve_size=True).type(dtype)
        gn = torch.Generator().manual_seed(123)
        mse = torch.nn.MSELoss()
        params = ref.common_utils.get_params('net', net, z0)
        opt = torch.optim.Adam(params, lr=lr)
        losses = []
        for i in range(iters):
            noise = torch.randn(z0.shape, generator=gn).type(dtype)
            opt.zero_grad()
            out = net(z0 + noise * sigma)
            loss = mse(ds(out), target)
            loss.backward()
            if i == 0:
                out0 = out.detach().clone()
                gnorm0 = np.array([p.grad.double().norm().item() for p in params])
                gsum0 = np.array([p.grad.double().sum().item() for p in params])
            losses.append(loss.item())
            opt.step()
        pnorm = np.array([p.detach().double().norm().item() for p in params])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), H=H, W=W, factor=factor, iters=iters, sigma=sigma, lr=lr,
                        losses=np.array(losses), out0=out0.numpy(), gnorm0=gnorm0, gsum0=gsum0, pnorm=pnorm,
                        dtype=str(dtype))
    print(name, 'losses', losses)


def _bn_state(net):
    """running_mean / running_var / num_batches_tracked of every BatchNorm, in state_dict order, concatenated."""
    sd = net.state_dict()
    rm = np.concatenate([sd[k].numpy().ravel() for k in sd if k.endswith('running_mean')])
    rv = np.concatenate([sd[k].numpy().ravel() for k in sd if k.endswith('running_var')])
    nbt = np.array([float(sd[k]) for k in sd if k.endswith('num_batches_tracked')])
    return rm.copy(), rv.copy(), nbt


def run_baseline_case(kind, threads=8, iters=2):
    """One-/two-step fixtures of the UNMODIFIED reference at the BASELINE.json shapes, on the reference's own data
    (committed copies under tests/golden/data): 'denoise512' (denoising.ipynb c4-c10, F16 512x512, sigma 25),
    'inpaint512' (inpainting.ipynb c5-c17, kate 512x512 + mask, skip=128, nearest), 'sr_zebra' (super-resolution.ipynb
    c5-c10, zebra 384x576 -> 96x144, Lanczos-2 x4), 'sr1024' (BASELINE wording 256 -> 1024: synthetic LR target).
    Full tensors at these sizes are too big for fixtures: the network output is stored on a stride-4 grid together with
    its first two moments and the loss; gradients as per-tensor norms / sums plus slices."""
    import importlib
    torch.set_num_threads(threads)
    dtype = torch.float32
    data = os.path.join(HERE, 'data')
    with ref_harness.reference_modules() as ref:
        cu, models = ref.common_utils, ref.models
        mask = None
        ds = None
        extra = {}
        if kind == 'denoise512':
            img_pil = cu.crop_image(cu.get_image(os.path.join(data, 'F16_GT.png'), -1)[0], d=32)
            img_np = cu.pil_to_np(img_pil)
            np.random.seed(0)
            _, noisy_np = ref.denoising_utils.get_noisy_image(img_np, 25 / 255.)
            target = cu.np_to_torch(noisy_np).type(dtype)
            H, W = img_np.shape[1:]
            sigma = 1. / 30
            torch.manual_seed(0)
            net = models.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                                 upsample_mode='bilinear').type(dtype)
            extra['target_sum'] = float(target.double().sum())
        elif kind == 'inpaint512':
            img_pil, _ = cu.get_image(os.path.join(data, 'kate.png'), -1)
            mask_pil, _ = cu.get_image(os.path.join(data, 'kate_mask.png'), -1)
            mask_pil = cu.crop_image(mask_pil, 64)
            img_pil = cu.crop_image(img_pil, 64)
            img_np, mask_np = cu.pil_to_np(img_pil), cu.pil_to_np(mask_pil)
            target = cu.np_to_torch(img_np).type(dtype)
            mask = cu.np_to_torch(mask_np).type(dtype)
            H, W = img_np.shape[1:]
            sigma = 0.03
            torch.manual_seed(0)
            net = models.skip(32, img_np.shape[0], num_channels_down=[128] * 5, num_channels_up=[128] * 5,
                              num_channels_skip=[128] * 5, filter_size_up=3, filter_size_down=3, upsample_mode='nearest',
                              filter_skip_size=1, need_sigmoid=True, need_bias=True, pad='reflection',
                              act_fun='LeakyReLU').type(dtype)
            extra['mask_sum'] = float(mask.double().sum())
        elif kind in ('sr_zebra', 'sr1024'):
            sru = importlib.import_module('utils.sr_utils')
            dsm = importlib.import_module('models.downsampler')
            if kind == 'sr_zebra':
                imgs = sru.load_LR_HR_imgs_sr(os.path.join(data, 'zebra_GT.png'), -1, 4, 'CROP')
                target = cu.np_to_torch(imgs['LR_np']).type(dtype)
                H, W = imgs['HR_pil'].size[1], imgs['HR_pil'].size[0]
            else:
                H = W = 1024
                g = torch.Generator().manual_seed(2)
                target = torch.rand(1, 3, H // 4, W // 4, generator=g).type(dtype)
            sigma = 0.03
            torch.manual_seed(0)
            net = models.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                                 upsample_mode='bilinear').type(dtype)
            ds = dsm.Downsampler(n_planes=3, factor=4, kernel_type='lanczos2', phase=0.5, preserve_size=True).type(dtype)
            extra['target_sum'] = float(target.double().sum())
        else:
            raise ValueError(kind)
        torch.manual_seed(1)
        z0 = cu.get_noise(32, 'noise', (H, W)).type(dtype).detach()
        gn = torch.Generator().manual_seed(123)
        mse = torch.nn.MSELoss()
        params = cu.get_params('net', net, z0)
        opt = torch.optim.Adam(params, lr=0.01)
        losses = []
        for i in range(iters):
            noise = torch.randn(z0.shape, generator=gn).type(dtype)
            opt.zero_grad()
            out = net(z0 + noise * sigma)
            o = ds(out) if ds is not None else out
            loss = mse(o * mask, target * mask) if mask is not None else mse(o, target)
            loss.backward()
            if i == 0:
                out0 = out.detach().clone()
                gnorm0 = np.array([p.grad.double().norm().item() for p in params])
                gsum0 = np.array([p.grad.double().sum().item() for p in params])
                g_head_w = params[-2].grad.detach().clone().numpy()
                g_up0_w_slice = params[-10].grad.detach()[:4, :8].clone().numpy()      # L0.up.w[:4,:8]
                g_d2_4_slice = params[4 * 12 + 8].grad.detach()[:4, :8].clone().numpy()  # L4.d2.w[:4,:8] (deepest level)
                g_skip0_w = params[0].grad.detach().clone().numpy()
                rm1, rv1, nbt1 = _bn_state(net)
            losses.append(loss.item())
            opt.step()
        pnorm = np.array([p.detach().double().norm().item() for p in params])
        rm, rv, nbt = _bn_state(net)
    np.savez_compressed(os.path.join(HERE, 'baseline_' + kind + '_fp32.npz'), H=H, W=W, iters=iters, sigma=sigma, lr=0.01,
                        losses=np.array(losses), out0_sub=out0.numpy()[:, :, ::4, ::4].astype(np.float32),
                        out0_mean=out0.double().mean().item(), out0_sq=(out0.double() ** 2).mean().item(),
                        gnorm0=gnorm0, gsum0=gsum0, g_head_w=g_head_w, g_up0_w_slice=g_up0_w_slice,
                        g_d2_4_slice=g_d2_4_slice, g_skip0_w=g_skip0_w, pnorm=pnorm, rm1=rm1, rv1=rv1, nbt1=nbt1, rm=rm,
                        rv=rv, nbt=nbt, **extra)
    print(kind, H, W, 'losses', losses)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'baseline':   # BASELINE-shape fixtures on the reference's data
        for kind in (sys.argv[2:] or ['denoise512', 'inpaint512', 'sr_zebra', 'sr1024']):
            run_baseline_case(kind, threads=int(os.environ.get('DIP_GOLD_THREADS', '8')))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'inpaint':   # only the inpainting (skip=128) fixtures
        run_case('inpaint64x96_nearest_masked_skip128_fp64', 64, 96, 'nearest', 3, sigma=0.03, masked=True,
                 dtype=torch.float64, skip_n11=128)
        run_case('inpaint64x96_nearest_masked_skip128_fp32', 64, 96, 'nearest', 3, sigma=0.03, masked=True,
                 dtype=torch.float32, skip_n11=128)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'variants':   # only the flash-no-flash / restoration fixtures
        for dt, tag in ((torch.float64, 'fp64'), (torch.float32, 'fp32')):
            run_variant('flash64x96_in3_mixed_' + tag, 64, 96, 3, 3, FLASH_MODES, dtype=dt)
            run_variant('restore64_out1_masked_' + tag, 64, 64, 32, 1, 'bilinear', masked=True, dtype=dt)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'vase':   # inpainting.ipynb "vase": num_channels_skip = 0, meshgrid input, nearest
        for dt, tag in ((torch.float64, 'fp64'), (torch.float32, 'fp32')):
            run_variant('vase64x96_in2_skip0_masked_' + tag, 64, 96, 2, 3, 'nearest', masked=True, dtype=dt, skip_ch=0, meshgrid=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'snail':   # denoising.ipynb c8:13-23 "snail": per-scale widths and skips, 3-channel noise input
        for dt, tag in ((torch.float64, 'fp64'), (torch.float32, 'fp32')):
            run_variant('snail64x96_in3_w8to128_' + tag, 64, 96, 3, 3, 'bilinear', sigma=1. / 30, dtype=dt,
                        chans=[8, 16, 32, 64, 128], skips=[0, 0, 0, 4, 4])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'kate_restore':   # restoration.ipynb c7:28-36 kate: widths 16..128, no skips, 'avg' downsampling, masked
        for dt, tag in ((torch.float64, 'fp64'), (torch.float32, 'fp32')):
            run_variant('restorekate64x96_avg_w16to128_' + tag, 64, 96, 32, 3, 'bilinear', sigma=0.0, masked=True, dtype=dt,
                        chans=[16, 32, 64, 128, 128], skips=[0, 0, 0, 0, 0], downsample_mode='avg')
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'sr':   # only the super-resolution fixtures
        run_downsampler_cases()
        run_sr_case('sr64x96_fp64', 64, 96, 3, torch.float64)
        run_sr_case('sr64x96_fp32', 64, 96, 3, torch.float32)
        sys.exit(0)
    run_case('denoise64_bilinear_fp32', 64, 64, 'bilinear', 4)
    run_case('denoise64_bilinear_fp64', 64, 64, 'bilinear', 4, dtype=torch.float64)
    run_case('denoise96x64_nearest_masked_fp64', 96, 64, 'nearest', 3, sigma=0.03, masked=True, dtype=torch.float64)
    run_case('inpaint64x96_nearest_masked_skip128_fp64', 64, 96, 'nearest', 3, sigma=0.03, masked=True,
             dtype=torch.float64, skip_n11=128)
    run_case('inpaint64x96_nearest_masked_skip128_fp32', 64, 96, 'nearest', 3, sigma=0.03, masked=True,
             dtype=torch.float32, skip_n11=128)
    run_downsampler_cases()
    run_sr_case('sr64x96_fp64', 64, 96, 3, torch.float64)
    run_sr_case('sr64x96_fp32', 64, 96, 3, torch.float32)
    for dt, tag in ((torch.float64, 'fp64'), (torch.float32, 'fp32')):
        run_variant('flash64x96_in3_mixed_' + tag, 64, 96, 3, 3, FLASH_MODES, dtype=dt)
        run_variant('restore64_out1_masked_' + tag, 64, 64, 32, 1, 'bilinear', masked=True, dtype=dt)
