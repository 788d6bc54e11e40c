"""P4 full-budget fixture: the UNMODIFIED reference (imported from /root/reference) on the denoising.ipynb F16 problem,
torch-CPU fp32, BASELINE.json's budget of 2000 iterations, hyper-parameters of denoising.ipynb c8 (reg_noise_std 1/30,
LR 0.01, adam, exp_weight 0.99), closure of denoising.ipynb c10 (EMA out_avg, PSNR_noisy / PSNR_gt / PSNR_gt_sm,
backtracking).  The only deviations from the notebook, all needed to make two implementations comparable:
  * seeds (the reference seeds nothing): np.random.seed(0) before get_noisy_image, torch.manual_seed(0) before get_net,
    torch.manual_seed(1) before get_noise;
  * the per-iteration perturbation is drawn from a dedicated torch.Generator (seed 123) so that the engine run
    (tests/test_full_budget_gpu.py) can consume the identical stream;
  * dtype = torch.FloatTensor (CPU), so `.cuda()` in the backtracking branch becomes `.type(dtype)`.

Run twice with different thread counts: the difference between the two files is the reference's own run-to-run spread.
  python tests/golden/make_f16_full.py --threads 4 --out tests/golden/f16_full_t4.npz
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, default=4)
ap.add_argument("--iters", type=int, default=2000)
ap.add_argument("--out", required=True)
args = ap.parse_args()
torch.set_num_threads(args.threads)

with ref_harness.reference_modules() as ref:
    cu, du, models = ref.common_utils, ref.denoising_utils, ref.models
    from skimage.measure import compare_psnr
    dtype = torch.FloatTensor
    sigma_ = 25 / 255.
    img_pil = cu.crop_image(cu.get_image(os.path.join(HERE, "data", "F16_GT.png"), -1)[0], d=32)
    img_np = cu.pil_to_np(img_pil)
    np.random.seed(0)
    img_noisy_pil, img_noisy_np = du.get_noisy_image(img_np, sigma_)
    reg_noise_std, LR, exp_weight, show_every = 1. / 30., 0.01, 0.99, 100
    torch.manual_seed(0)
    net = models.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                         upsample_mode='bilinear').type(dtype)
    torch.manual_seed(1)
    net_input = cu.get_noise(32, 'noise', (img_pil.size[1], img_pil.size[0])).type(dtype).detach()
    mse = torch.nn.MSELoss().type(dtype)
    img_noisy_torch = cu.np_to_torch(img_noisy_np).type(dtype)
    net_input_saved = net_input.detach().clone()
    noise = net_input.detach().clone()
    gen = torch.Generator().manual_seed(123)
    st = dict(i=0, out_avg=None, last_net=None, psrn_noisy_last=0, fallbacks=0)
    rec = dict(loss=[], psnr_noisy=[], psnr_gt=[], psnr_gt_sm=[])
    t0 = time.time()

    def closure():
        global net_input
        net_input = net_input_saved + (noise.normal_(generator=gen) * reg_noise_std)
        out = net(net_input)
        if st['out_avg'] is None:
            st['out_avg'] = out.detach()
        else:
            st['out_avg'] = st['out_avg'] * exp_weight + out.detach() * (1 - exp_weight)
        total_loss = mse(out, img_noisy_torch)
        total_loss.backward()
        o = out.detach().cpu().numpy()[0]
        psrn_noisy = compare_psnr(img_noisy_np, o)
        psrn_gt = compare_psnr(img_np, o)
        psrn_gt_sm = compare_psnr(img_np, st['out_avg'].detach().cpu().numpy()[0])
        rec['loss'].append(total_loss.item()); rec['psnr_noisy'].append(psrn_noisy)
        rec['psnr_gt'].append(psrn_gt); rec['psnr_gt_sm'].append(psrn_gt_sm)
        if st['i'] % 50 == 0:
            print('it %05d loss %f noisy %f gt %f gt_sm %f  (%.0f s)' % (st['i'], total_loss.item(), psrn_noisy, psrn_gt,
                                                                       psrn_gt_sm, time.time() - t0), flush=True)
        if st['i'] % show_every:
            if psrn_noisy - st['psrn_noisy_last'] < -5:
                st['fallbacks'] += 1
                for new_param, net_param in zip(st['last_net'], net.parameters()):
                    net_param.data.copy_(new_param.type(dtype))
                return total_loss * 0
            else:
                st['last_net'] = [x.detach().cpu() for x in net.parameters()]
                st['psrn_noisy_last'] = psrn_noisy
        st['i'] += 1
        return total_loss

    p = cu.get_params('net', net, net_input)
    cu.optimize('adam', p, closure, LR, args.iters)
    out_np = cu.torch_to_np(net(net_input))
    np.savez_compressed(args.out, threads=args.threads, iters=args.iters, fallbacks=st['fallbacks'],
                        loss=np.array(rec['loss']), psnr_noisy=np.array(rec['psnr_noisy']), psnr_gt=np.array(rec['psnr_gt']),
                        psnr_gt_sm=np.array(rec['psnr_gt_sm']), final_psnr_gt=compare_psnr(img_np, out_np),
                        seconds=time.time() - t0,
                        out_avg_u8=np.clip(np.round(st['out_avg'].numpy()[0] * 255), 0, 255).astype(np.uint8))
    print("done", time.time() - t0, "s; final psnr_gt", rec['psnr_gt'][-1], "gt_sm", rec['psnr_gt_sm'][-1], "fallbacks", st['fallbacks'])
