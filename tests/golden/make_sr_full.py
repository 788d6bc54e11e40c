"""P4 full-budget fixture for BASELINE config 3: the UNMODIFIED reference (imported from /root/reference) on the
super-resolution.ipynb zebra problem (x4, 576x384 -> 144x96, Lanczos-2 downsampler in the loss), torch-CPU fp32, the notebook's
budget of 2000 iterations and hyper-parameters (c7: reg_noise_std 0.03, LR 0.01, adam; c8: get_net(32, 'skip', 'reflection',
128, 128, 4, 5, 'bilinear'), Downsampler(3, 4, 'lanczos2', 0.5, preserve_size=True)), closure of c10 (PSNR_LR / PSNR_HR history).
Deviations from the notebook (as in make_f16_full.py): seeds (torch.manual_seed(0) before get_net, 1 before get_noise) and the
per-iteration perturbation from a dedicated torch.Generator (seed 123) that the engine run consumes as well.
  python tests/golden/make_sr_full.py --threads 4 --out tests/golden/sr_full_t4.npz
"""
import argparse
import importlib
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
    cu, models = ref.common_utils, ref.models
    sru = importlib.import_module("utils.sr_utils")
    dsm = importlib.import_module("models.downsampler")
    from skimage.measure import compare_psnr
    dtype = torch.FloatTensor
    factor = 4
    imgs = sru.load_LR_HR_imgs_sr(os.path.join(HERE, "data", "zebra_GT.png"), -1, factor, "CROP")
    reg_noise_std, LR = 0.03, 0.01
    torch.manual_seed(1)
    net_input = cu.get_noise(32, "noise", (imgs["HR_pil"].size[1], imgs["HR_pil"].size[0])).type(dtype).detach()
    torch.manual_seed(0)
    net = models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                         upsample_mode="bilinear").type(dtype)
    mse = torch.nn.MSELoss().type(dtype)
    img_LR_var = cu.np_to_torch(imgs["LR_np"]).type(dtype)
    downsampler = dsm.Downsampler(n_planes=3, factor=factor, kernel_type="lanczos2", phase=0.5, preserve_size=True).type(dtype)
    net_input_saved = net_input.detach().clone()
    noise = net_input.detach().clone()
    gen = torch.Generator().manual_seed(123)
    rec = dict(loss=[], psnr_LR=[], psnr_HR=[])
    st = dict(i=0)
    t0 = time.time()

    def closure():
        global net_input
        net_input = net_input_saved + (noise.normal_(generator=gen) * reg_noise_std)
        out_HR = net(net_input)
        out_LR = downsampler(out_HR)
        total_loss = mse(out_LR, img_LR_var)
        total_loss.backward()
        psnr_LR = compare_psnr(imgs["LR_np"], cu.torch_to_np(out_LR))
        psnr_HR = compare_psnr(imgs["HR_np"], cu.torch_to_np(out_HR))
        rec["loss"].append(total_loss.item()); rec["psnr_LR"].append(psnr_LR); rec["psnr_HR"].append(psnr_HR)
        if st["i"] % 50 == 0:
            print("it %05d loss %f PSNR_LR %.3f PSNR_HR %.3f (%.0f s)" % (st["i"], total_loss.item(), psnr_LR, psnr_HR, time.time() - t0),
                  flush=True)
        st["i"] += 1
        return total_loss

    p = cu.get_params("net", net, net_input)
    cu.optimize("adam", p, closure, LR, args.iters)
    np.savez_compressed(args.out, threads=args.threads, iters=args.iters, loss=np.array(rec["loss"]),
                        psnr_LR=np.array(rec["psnr_LR"]), psnr_HR=np.array(rec["psnr_HR"]), seconds=time.time() - t0)
    print("done", time.time() - t0, "s; final PSNR_LR", rec["psnr_LR"][-1], "PSNR_HR", rec["psnr_HR"][-1])
