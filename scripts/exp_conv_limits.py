"""Where does tc_conv_kernel spend its time?  L0 up-conv shape (512x512, 132->128, 3x3), timing with parts disabled."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import dip_engine as de
H = W = 512
a = torch.randn(H + 2, W + 2, 132, device="cuda")
w = torch.randn(128, 132, 3, 3, device="cuda") * 0.03
b = torch.randn(128, device="cuda")
a1 = torch.randn(H, W, 128, device="cuda")
w1 = torch.randn(128, 128, 1, 1, device="cuda") * 0.1
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for patch in ("1", "0"):
    if patch == "0": os.environ["DIP_NO_PATCH"] = "1"
    else: os.environ.pop("DIP_NO_PATCH", None)
    for flags, name in ((0, "baseline"), (4, "no epilogue"), (1, "no A loads"), (2, "no B loads"), (3, "no loads"), (8, "no MMAs"), (12, "no MMA, no epilogue"), (7, "only MMAs")):
        os.environ["DIP_DBG_FLAGS"] = str(flags)
        stats = torch.zeros(256 * 16, dtype=torch.float64, device="cuda")
        t3 = timeit(lambda: de.op_conv_fprop(a, w, b, 3, 1, 0, 0, H, W, rot=4, stats=stats))
        t3n = timeit(lambda: de.op_conv_fprop(a, w, b, 3, 1, 0, 0, H, W, rot=4, stats=None))
        t1 = timeit(lambda: de.op_conv_fprop(a1, w1, b, 1, 1, 0, 0, H, W, stats=stats))
        t1n = timeit(lambda: de.op_conv_fprop(a1, w1, b, 1, 1, 0, 0, H, W, stats=None))
        print("patch=%s %-22s 3x3: %7.1f us (no stats %7.1f)   1x1: %7.1f us (no stats %7.1f)" % (patch, name, t3, t3n, t1, t1n))
