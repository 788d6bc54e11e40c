"""Executes the code cells of a reference notebook UNCHANGED -- test infrastructure.

jupyter / nbformat are not installed: the .ipynb is JSON; every code cell's source is compiled as is, except that
IPython-only lines (`%magic`, `!shell`) are dropped, and exec'd in ONE namespace, like a kernel would.  `overrides`
(e.g. num_iter, PLOT) are re-applied after every cell to names the notebook has already defined, which is how a user
would edit the config cell's values -- the cell sources themselves are never edited.

`models` / `utils` resolve to whatever sys.path says (tests/conftest.py puts deep-image-prior_b200/ first), the three
environment shims of SURVEY.md section 4 (matplotlib stub, PIL.Image.ANTIALIAS, skimage.measure.compare_psnr) come from
oracle/ref_harness.py, and the working directory is the notebook's own directory so that its relative data paths work.
"""
import contextlib
import io
import json
import os


def code_cells(path):
    nb = json.load(open(path))
    for idx, cell in enumerate(nb["cells"]):
        if cell["cell_type"] != "code":
            continue
        src = "".join(cell["source"])
        kept = [ln for ln in src.split("\n") if not ln.lstrip().startswith(("%", "!"))]
        yield idx, "\n".join(kept), src


def run_notebook(path, overrides=None, stop_after=None, quiet=True):
    """Returns the namespace after the last executed cell.  stop_after: JSON cell index after which to stop."""
    from oracle import ref_harness
    ref_harness._install_shims()
    overrides = dict(overrides or {})
    ns = {"__name__": "__main__"}
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(path)))
    sink = io.StringIO()
    try:
        for idx, code, _ in code_cells(path):
            obj = compile(code, "%s:c%d" % (os.path.basename(path), idx), "exec")
            with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
                exec(obj, ns)
            for k, v in overrides.items():
                if k in ns:
                    ns[k] = v
            if stop_after is not None and idx >= stop_after:
                break
    finally:
        os.chdir(cwd)
    ns["__stdout__"] = sink.getvalue()
    return ns
