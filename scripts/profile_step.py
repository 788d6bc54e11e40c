"""Runs N iterations of the device runner (for ncu launch lists / captures). usage: profile_step.py [iters] [H] [W]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
from oracle import dip_oracle as O
import dip_engine as de
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
W = int(sys.argv[3]) if len(sys.argv) > 3 else 512
cfg = O.SkipConfig(upsample_mode="bilinear")
params = [p.detach().cuda() for p in O.init_params(cfg, seed=0)]
grads = [torch.zeros_like(p) for p in params]
plan = de.Plan(32, 3, 5, 128, 4, True, H, W)
plan.bind(params, grads)
for p, g in zip(params, grads):
    p.grad = g
adam = de.FusedAdam(params, lr=0.01)
adam._bind(grads)
z0 = torch.rand(1, 32, H, W, device="cuda") * 0.1
target = torch.rand(1, 3, H, W, device="cuda")
out = torch.empty(1, 3, H, W, device="cuda")
de.run_iterations(plan, adam, z0, target, None, 1 / 30., 1, iters, 0.01, out=out)
torch.cuda.synchronize()
print("done", plan.num_launches())
