"""How long does one tcgen05.mma kind::tf32 (M128 N128 K8) take?  Only-MMA mode, n MMAs per k-block."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import dip_engine as de
H = W = 512
a = torch.randn(H + 2, W + 2, 128, device="cuda")
w = torch.randn(128, 128, 3, 3, device="cuda") * 0.03
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
os.environ.pop("DIP_DBG_STAGES", None)
for flags, name in ((7, "wait+fence+mma+commit"), (7 + 32, "no fence"), (7 + 64 + 128, "no wait, no commit (fence+mma only)"), (7 + 32 + 64 + 128, "mma only")):
    os.environ["DIP_DBG_FLAGS"] = str(flags)
    for n in (1, 4, 8):
        os.environ["DIP_DBG_NMMA"] = str(n)
        t = timeit(lambda: de.op_conv_fprop(a, w, None, 3, 1, 0, 0, H, W))
        kblocks = 2048 * 36 / 148
        print("%-40s MMAs per k-block %2d: %7.1f us  -> %6.0f cycles/k-block @1.9GHz" % (name, n, t, t * 1e-6 * 1.9e9 / kblocks))
