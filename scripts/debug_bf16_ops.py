"""Prints the error of every bf16 single-op case (tests/test_conv_ops_gpu.py, precision 2) instead of stopping at the first
failure: python scripts/debug_bf16_ops.py fprop|dgrad|wgrad|s2"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "deep-image-prior_b200"))
import torch
import test_conv_ops_gpu as T

what = sys.argv[1]
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 2
def run(fn, *a):
    try:
        fn(*a)
        print("ok  ", fn.__name__, a, flush=True)
    except AssertionError as e:
        print("FAIL", fn.__name__, a, str(e).split("\n")[0][:200], flush=True)
    except Exception as e:
        print("ERR ", fn.__name__, a, repr(e)[:300], flush=True)
if what == "fprop":
    for c in T.CASES: run(T.test_fprop, c, prec)
elif what == "dgrad":
    for c in [(128, 3, 32, 32, 0), (132, 3, 32, 32, 4), (128, 1, 32, 32, 0), (128, 3, 10, 20, 0), (132, 3, 2, 2, 4),
              (128, 3, 64, 128, 0), (128, 3, 254, 254, 0), (132, 3, 200, 312, 4), (128, 3, 268, 148, 0)]: run(T.test_dgrad, c, prec)
elif what == "wgrad":
    for c in T.CASES: run(T.test_wgrad, c, prec)
elif what == "s2":
    for c in [(128, 16, 16), (128, 64, 64), (128, 129, 128), (32, 40, 24), (128, 9, 13), (128, 2, 2)]: run(T.test_dgrad_stride2_phases, c, prec)
