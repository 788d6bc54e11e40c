"""Recipe for oracle/_ref/: the UNMODIFIED reference, taken from the sources where they lie under /root/reference
-- TEST INFRASTRUCTURE ONLY (git-ignored, never committed; it travels to the GPU box with the working tree like the
built libdip.so does).  Run by __graft_entry__.build() whenever /root/reference is present, or by hand:

    python oracle/make_ref.py

What lands in oracle/_ref/ and who uses it:
  models/ utils/     the reference's Python packages  -> bench.py --impl reference / cpu_baseline (kind "reference":
                     the real thing timed on the GPU box's host cores instead of the oracle port)
  *.ipynb            the task notebooks               -> tests/test_notebooks_gpu.py executes their code cells
                     UNCHANGED against this repo's `models` / `utils` (north_star: "existing notebooks run unchanged")
  data/              the images the notebooks load by relative path
Nothing under deep-image-prior_b200/ may import from here.
"""
import os
import shutil
import sys

REF = os.environ.get("DIP_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

NOTEBOOKS = ["denoising.ipynb", "super-resolution.ipynb", "inpainting.ipynb", "restoration.ipynb", "flash-no-flash.ipynb",
             "sr_prior_effect.ipynb"]
DATA_DIRS = ["denoising", "sr", "inpainting", "restoration", "flash_no_flash"]


def make(verbose=True):
    if not os.path.isdir(os.path.join(REF, "models")):
        if verbose:
            print("oracle/make_ref.py: %s not present -- keeping whatever oracle/_ref already holds" % REF)
        return False
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for pkg in ("models", "utils"):
        shutil.copytree(os.path.join(REF, pkg), os.path.join(OUT, pkg), ignore=ignore)
    for nb in NOTEBOOKS:
        shutil.copy(os.path.join(REF, nb), os.path.join(OUT, nb))
    for d in DATA_DIRS:
        shutil.copytree(os.path.join(REF, "data", d), os.path.join(OUT, "data", d))
    for root, dirs, files in os.walk(OUT):          # the checkout is read-only; the copy must be removable
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write("verbatim copy of %s {models,utils,%s,data/{%s}} made by oracle/make_ref.py\n"
                % (REF, ",".join(NOTEBOOKS), ",".join(DATA_DIRS)))
    if verbose:
        print("oracle/_ref populated from", REF)
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
