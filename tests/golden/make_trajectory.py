"""P3/P4 fixture: loss / PSNR trajectory of the oracle (= the reference's graph on torch-CPU fp32) on a 256x256 synthetic
denoising problem, 400 iterations, fixed seeds and a fixed per-iteration noise stream.  Run twice with different thread
counts to record the reference's own run-to-run spread.  python tests/golden/make_trajectory.py"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dip_oracle as O

H = W = 256
ITERS = 400
CHECK = [0, 1, 2, 5, 10, 20, 50, 100, 200, 300, 399]


def problem():
    g = torch.Generator().manual_seed(2024)
    clean = torch.rand(1, 3, H // 16, W // 16, generator=g)
    clean = torch.nn.functional.interpolate(clean, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    noisy = (clean + torch.randn(clean.shape, generator=g) * (25. / 255.)).clamp(0, 1)
    return clean, noisy


def run(threads):
    torch.set_num_threads(threads)
    cfg = O.SkipConfig(upsample_mode="bilinear")
    params = O.init_params(cfg, seed=0)
    z0 = O.get_noise(32, (H, W), seed=1)
    clean, noisy = problem()
    gn = torch.Generator().manual_seed(123)
    noises = [torch.randn(z0.shape, generator=gn) for _ in range(ITERS)]
    rec = {"loss": [], "psnr_gt": []}

    def record(i, out, loss, grads):
        rec["loss"].append(loss)
        rec["psnr_gt"].append(O.psnr(clean.numpy()[0], out.numpy()[0]))
    t = time.time()
    O.run(cfg, params, z0, noisy, noises, 1. / 30, 0.01, record=record)
    print("threads", threads, "time", time.time() - t, [round(rec["psnr_gt"][i], 3) for i in CHECK])
    return rec


if __name__ == "__main__":
    a = run(8)
    b = run(4)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectory256.npz"), H=H, W=W, iters=ITERS,
                        loss_a=np.array(a["loss"]), psnr_a=np.array(a["psnr_gt"]), loss_b=np.array(b["loss"]),
                        psnr_b=np.array(b["psnr_gt"]), check=np.array(CHECK))
