import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import models
torch.manual_seed(0)
net = models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode="bilinear").type(torch.cuda.FloatTensor)
z = torch.rand(1, 32, 64, 64).cuda() * 0.1
out = net(z)
torch.cuda.synchronize()
sd = net.state_dict()
for k, v in sd.items():
    if "num_batches" in k:
        print(k, int(v), float(sd[k.replace("num_batches_tracked", "running_mean")].abs().mean()), float(sd[k.replace("num_batches_tracked", "running_var")].mean()))
import dip_engine as de
plan = list(net._dip_plans.values())[0]
print("launches", plan.num_launches())
