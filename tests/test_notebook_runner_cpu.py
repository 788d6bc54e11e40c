"""CPU checks of the notebook-cell runner (the GPU tests in tests/test_notebooks_gpu.py execute the cells)."""
import os

import pytest

from notebook_runner import code_cells

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
NOTEBOOKS = ["denoising.ipynb", "super-resolution.ipynb", "inpainting.ipynb", "flash-no-flash.ipynb", "restoration.ipynb"]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "denoising.ipynb")), reason="oracle/_ref not populated")
@pytest.mark.parametrize("name", NOTEBOOKS)
def test_cells_compile_and_only_ipython_lines_are_dropped(name):
    n = 0
    for idx, code, src in code_cells(os.path.join(REF, name)):
        compile(code, "%s:c%d" % (name, idx), "exec")
        dropped = [ln for ln in src.split("\n") if ln not in code.split("\n")]
        assert all(ln.lstrip().startswith(("%", "!")) for ln in dropped), dropped
        n += 1
    assert n >= 6
