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


def run_case(name, H, W, mode, iters, sigma=1. / 30, lr=0.01, masked=False, dtype=torch.float32, threads=8):
    torch.set_num_threads(threads)
    with ref_harness.reference_modules() as ref:
        torch.manual_seed(0)
        net = ref.models.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
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
                        dtype=str(dtype), state_keys=np.array(keys))
    print(name, 'losses', losses)


if __name__ == '__main__':
    run_case('denoise64_bilinear_fp32', 64, 64, 'bilinear', 4)
    run_case('denoise64_bilinear_fp64', 64, 64, 'bilinear', 4, dtype=torch.float64)
    run_case('denoise96x64_nearest_masked_fp64', 96, 64, 'nearest', 3, sigma=0.03, masked=True, dtype=torch.float64)
