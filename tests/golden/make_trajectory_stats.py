"""P4 fixture (distribution-level parity): final PSNR of the reference's graph (oracle on torch-CPU fp32) on the 256x256
synthetic denoising problem of make_trajectory.py, 400 iterations, for several per-iteration noise streams.  The DIP
trajectory is chaotic, so single runs are not comparable to 1e-3 dB; the mean over noise streams is.
python tests/golden/make_trajectory_stats.py [first_seed] [n_seeds] [threads]  ->  trajectory_stats256_<first>.npz"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dip_oracle as O
from make_trajectory import H, W, ITERS, problem

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.set_num_threads(threads)
cfg = O.SkipConfig(upsample_mode="bilinear")
clean, noisy = problem()
seeds, finals, tails, lossf = [], [], [], []
for sd in range(first, first + n):
    params = O.init_params(cfg, seed=0)
    z0 = O.get_noise(32, (H, W), seed=1)
    gn = torch.Generator().manual_seed(5000 + sd)
    noises = [torch.randn(z0.shape, generator=gn) for _ in range(ITERS)]
    ps, ls = [], []

    def record(i, out, loss, grads):
        ps.append(O.psnr(clean.numpy()[0], out.numpy()[0]))
        ls.append(loss)
    t = time.time()
    O.run(cfg, params, z0, noisy, noises, 1. / 30, 0.01, record=record)
    seeds.append(sd); finals.append(ps[-1]); tails.append(float(np.mean(ps[-50:]))); lossf.append(float(np.mean(ls[-50:])))
    print("seed", sd, "time %.0f s" % (time.time() - t), "final %.3f tail-mean %.3f" % (ps[-1], tails[-1]), flush=True)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectory_stats256_%d.npz" % first),
                    seeds=np.array(seeds), final=np.array(finals), tail_mean=np.array(tails), tail_loss=np.array(lossf),
                    iters=ITERS, threads=threads)
