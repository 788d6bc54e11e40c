"""GPU debugging aid: per-buffer / per-parameter comparison of the engine against the CPU oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
from oracle import dip_oracle as O  # noqa: E402
import dip_engine as de  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main(H=64, W=64, mode="bilinear", prec="fp32"):
    cfg = O.SkipConfig(upsample_mode=mode)
    params = O.init_params(cfg, seed=0)
    z0 = O.get_noise(32, (H, W), seed=1)
    g = torch.Generator().manual_seed(2)
    target = torch.rand(1, 3, H, W, generator=g)
    tape = {}
    out_ref = O.skip_forward(params, z0, cfg, tape=tape)
    for t in tape.values():
        t.retain_grad()
    loss = O.mse_loss(out_ref, target)
    loss.backward()
    dout = (2.0 * (out_ref.detach() - target) / out_ref.numel()).contiguous()
    plan = de.Plan(32, 3, 5, 128, 4, mode == "bilinear", H, W,
                   precision=de.PRECISION_TF32 if prec == "tf32" else de.PRECISION_FP32)
    dparams = [p.detach().cuda().contiguous() for p in params]
    dgrads = [torch.zeros_like(p) for p in dparams]
    plan.bind(dparams, dgrads)
    out = plan.forward(z0.cuda())
    plan.backward(dout.cuda())
    torch.cuda.synchronize()
    print("== %dx%d %s %s: out max abs err %.3e" % (H, W, mode, prec, (out.cpu() - out_ref.detach()).abs().max()))
    for l in range(5):
        for nm in ("raw_s", "raw_d1", "raw_d2", "raw_u", "raw_v"):
            ref = tape["L%d.%s" % (l, nm)]
            got = plan.buffer("L%d.%s" % (l, nm))
            gref = ref.grad[0].permute(1, 2, 0)
            ggot = plan.buffer("L%d.d%s%s" % (l, nm[0].upper(), nm[1:]))
            print("L%d.%-7s fwd %.2e   grad %.2e  (|g| %.2e)" % (l, nm, rel(got, ref[0].permute(1, 2, 0)), rel(ggot, gref), gref.norm()))
        cg = tape["L%d.cat" % l].grad[0].permute(1, 2, 0)
        cg = torch.cat([cg[:, :, 4:], cg[:, :, :4]], dim=2)
        print("L%d.dCat          grad %.2e" % (l, rel(plan.buffer("L%d.dCat" % l), cg)))
    names = [n for n, _ in O.param_layout(cfg)]
    for name, gd, p in zip(names, dgrads, params):
        e = rel(gd, p.grad)
        flag = " <<<" if e > 1e-2 and p.grad.norm() > 1e-7 else ""
        print("%-14s rel %.2e  |ref| %.2e |got| %.2e%s" % (name, e, p.grad.norm(), gd.norm(), flag))


if __name__ == "__main__":
    main(*( [int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]] if len(sys.argv) > 4 else []))
