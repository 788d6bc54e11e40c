"""Per-launch table of one optimisation step from the engine's own event timers (the numbers behind bench.py's rooflines):
class, algorithmic bytes / FLOPs, device time, achieved GB/s or TFLOP/s.   python scripts/hbm_breakdown.py [H W [skip]]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import torch
import bench
import dip_engine as de

H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W = int(sys.argv[2]) if len(sys.argv) > 2 else H
CS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
plan = de.Plan(32, 3, 5, 128, CS, True, H, W)
g = torch.Generator().manual_seed(0)
from oracle import dip_oracle as O
params = [p.detach().cuda().contiguous() for p in O.init_params(O.SkipConfig(skip_channels=CS), seed=0)]
grads = [torch.zeros_like(p) for p in params]
plan.bind(params, grads)
for p, gb in zip(params, grads):
    p.grad = gb
adam = de.FusedAdam(params, lr=0.01)
adam._bind(grads)
z0 = torch.rand(1, 32, H, W, device="cuda") * 0.1
target = torch.rand(1, 3, H, W, device="cuda")
de.run_iterations(plan, adam, z0, target, None, 1. / 30, 1, 5, 0.01)
os.environ["DIP_NO_SIDE"] = "1"
plan.set_timing(True)
N = 5
de.run_iterations(plan, adam, z0, target, None, 1. / 30, 1, N, 0.01)
torch.cuda.synchronize()
recs = plan.get_timing_records()
per = len(recs) // N
names = {0: "conv fprop", 1: "conv dgrad", 2: "wgrad"}
rows = []
for i in range(per):
    cls = recs[i][0]
    amt = recs[i][1]
    us = sorted(1000.0 * recs[i + k * per][2] for k in range(N))[N // 2]
    if cls < 16:
        nm, rate = names[cls], "%7.1f TF/s" % (amt / us / 1e6)
    else:
        nm = "%s<%d>" % (bench.HBM_NAMES[(cls - 16) // 8], (cls - 16) % 8)
        rate = "%7.0f GB/s" % (amt / us / 1e3)
    rows.append((nm, amt, us, rate))
tot = sum(r[2] for r in rows)
print("launches with a timer: %d, serialised %.1f us" % (per, tot))
for nm, amt, us, rate in rows:
    print("%-22s %10.2f M%s %8.1f us  %s" % (nm, amt / 1e6, "FLOP" if nm.startswith(("conv", "wgrad")) else "B   ", us, rate))
