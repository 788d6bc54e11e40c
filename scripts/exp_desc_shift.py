"""Experiment: can a K-major SWIZZLE_128B UMMA operand start at a row that is not a multiple of 8 (1024 B atom)?
Needed for reusing one shared-memory input patch for all 3x3 taps.  1x1 conv, tile = 128 consecutive pixels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import dip_engine as de
g = torch.Generator().manual_seed(0)
a = torch.randn(1, 256, 128, generator=g).cuda()          # [H=1][W=256][C=128] -> tiles of 128 consecutive pixels
w = (torch.randn(128, 128, 1, 1, generator=g) / 11.3).cuda()
ref = torch.einsum("hwc,nc->hwn", a.double(), w[:, :, 0, 0].double())
for shift in (0, 1, 2, 3, 4, 7, 8, 9):
    for bo in (0, 1):
        os.environ["DIP_DBG_SHIFT"] = str(shift); os.environ["DIP_DBG_BO"] = str(bo)
        d = de.op_conv_fprop(a, w, None, 1, 1, 0, 0, 1, 256).double()
        torch.cuda.synchronize()
        errs = []
        for y in range(1):
            for x0 in (0, 128):
                got = d[y, x0:x0 + 128 - shift]
                want = ref[y, x0 + shift:x0 + 128]
                errs.append(((got - want).norm() / want.norm()).item())
        print("shift %d base_offset %d: rel err %.3e" % (shift, bo, max(errs)))
