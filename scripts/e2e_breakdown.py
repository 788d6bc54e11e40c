"""Where does the notebook-path step (bench.py `e2e`) spend its time?  Host-side cost of each call of the closure-style
step, and the step rate with / without the per-step loss read-back.  (GPU box)"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import dip_engine as de
import models
H = W = 512
torch.manual_seed(0)
net = models.get_net(32, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                     upsample_mode="bilinear").type(torch.cuda.FloatTensor)
z0 = torch.rand(1, 32, H, W, device="cuda") * 0.1
target = torch.rand(1, 3, H, W, device="cuda")
params = list(net.parameters())
opt = de.FusedAdam(params, lr=0.01)
mse = torch.nn.MSELoss()
T = {}


def tick(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0)


def step(z, sync):
    t = time.perf_counter(); opt.zero_grad(); tick("zero_grad", t)
    t = time.perf_counter(); out = net(z); tick("net(z)", t)
    t = time.perf_counter(); loss = mse(out, target); tick("mse", t)
    t = time.perf_counter(); loss.backward(); tick("backward", t)
    t = time.perf_counter(); opt.step(); tick("opt.step", t)
    t = time.perf_counter()
    if sync:
        loss.item()
    tick("loss.item", t)


def run(n, sync, label):
    global T
    for _ in range(5):
        step(z0, sync)
    torch.cuda.synchronize()
    T = {}
    t0 = time.perf_counter()
    for _ in range(n):
        step(z0, sync)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("%-28s %.3f ms/step = %.1f it/s | host ms/step: %s" % (label, dt * 1e3, 1 / dt, ", ".join(
        "%s %.3f" % (k, v / n * 1e3) for k, v in T.items())))


run(60, True, "sync each step (loss.item)")
run(60, False, "no per-step sync")
os.environ["DIP_NO_GRAPH"] = "1"
run(60, True, "eager launches, sync")
run(60, False, "eager launches, no sync")
