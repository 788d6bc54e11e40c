"""Per-parameter gradient error vs an fp64 oracle: ours (fp32 / tf32 mode) next to torch-CUDA (cuDNN, TF32 on/off)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
from oracle import dip_oracle as O
import dip_engine as de

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 64)
cfg = O.SkipConfig(upsample_mode="bilinear")
p32 = O.init_params(cfg, seed=0)
z0 = O.get_noise(32, (H, W), seed=1)
g = torch.Generator().manual_seed(2)
target = torch.rand(1, 3, H, W, generator=g)

def oracle_grads(params, z, t):
    out = O.skip_forward(params, z, cfg)
    return out.detach(), torch.autograd.grad(O.mse_loss(out, t), params)

p64 = [p.detach().double().requires_grad_(True) for p in p32]
out64, g64 = oracle_grads(p64, z0.double(), target.double())
res = {}
for name, tf32 in (("cudnn_tf32", True), ("cuda_fp32", False)):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    pc = [p.detach().cuda().requires_grad_(True) for p in p32]
    o, gg = oracle_grads(pc, z0.cuda(), target.cuda())
    res[name] = (o.cpu(), [x.cpu() for x in gg])
for name, prec in (("ours_fp32", de.PRECISION_FP32), ("ours_tf32", de.PRECISION_TF32)):
    plan = de.Plan(32, 3, 5, 128, 4, True, H, W, precision=prec)
    dp = [p.detach().cuda().contiguous() for p in p32]
    dg = [torch.zeros_like(p) for p in dp]
    plan.bind(dp, dg)
    o = plan.forward(z0.cuda())
    dout = (2.0 * (o - target.cuda()) / o.numel()).contiguous()
    plan.backward(dout)
    torch.cuda.synchronize()
    res[name] = (o.cpu(), [x.cpu() for x in dg])
print("out max abs err vs fp64:", {k: float((v[0].double() - out64).abs().max()) for k, v in res.items()})
names = [n for n, _ in O.param_layout(cfg)]
gmax = max(x.norm().item() for x in g64)
print("%-14s %9s | %10s %10s %10s %10s" % ("param", "|g64|", "cudnn_tf32", "cuda_fp32", "ours_fp32", "ours_tf32"))
for i, n in enumerate(names):
    if g64[i].norm().item() < 1e-5 * gmax:
        continue
    print("%-14s %9.2e | %10.2e %10.2e %10.2e %10.2e" % (n, g64[i].norm(), rel(res["cudnn_tf32"][1][i], g64[i]),
          rel(res["cuda_fp32"][1][i], g64[i]), rel(res["ours_fp32"][1][i], g64[i]), rel(res["ours_tf32"][1][i], g64[i])))
