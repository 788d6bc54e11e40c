"""Import shims for running the UNMODIFIED reference (/root/reference) in this container -- TEST INFRASTRUCTURE ONLY.

Three shims, nothing else: (a) stub matplotlib (imported at utils/common_utils.py:11, not
installed); (b) PIL.Image.ANTIALIAS alias (removed in Pillow >= 10; used at utils/common_utils.py:110);
(c) skimage.measure.compare_psnr (skimage absent; old semantics: data_range 1 for non-negative float images).
/root/reference exists only in the build container, never on the GPU box: callers must handle `available() == False`.
"""
import importlib
import math
import os
import sys
import types

REF = os.environ.get("DIP_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def _install_shims():
    sys.dont_write_bytecode = True
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            for fn in ("figure", "imshow", "show", "plot", "subplot", "title"):
                setattr(plt, fn, lambda *a, **k: None)
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    from PIL import Image
    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    if "skimage" not in sys.modules:
        import numpy as np
        sk = types.ModuleType("skimage")
        me = types.ModuleType("skimage.measure")

        def compare_psnr(im_true, im_test, data_range=None):
            im_true = np.asarray(im_true, dtype=np.float64)
            im_test = np.asarray(im_test, dtype=np.float64)
            if data_range is None:
                data_range = 1.0 if im_true.min() >= 0 else 2.0
            return 10 * math.log10(data_range ** 2 / np.mean((im_true - im_test) ** 2))

        me.compare_psnr = compare_psnr
        sk.measure = me
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = me


class _RefModules:
    """Context manager: temporarily makes `models` / `utils` resolve to the reference's packages."""

    def __enter__(self):
        _install_shims()
        self.saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.") or
                      k == "utils" or k.startswith("utils.")}
        for k in self.saved:
            del sys.modules[k]
        sys.path.insert(0, REF)
        self.models = importlib.import_module("models")
        self.common_utils = importlib.import_module("utils.common_utils")
        self.denoising_utils = importlib.import_module("utils.denoising_utils")
        return self

    def __exit__(self, *exc):
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or
                  k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


def reference_modules():
    if not available():
        raise RuntimeError("reference checkout not present at " + REF)
    return _RefModules()
