"""Host-side task utilities of the super-resolution / inpainting notebooks (callers of the hot path): same names and
results as the reference's utils/sr_utils.py and utils/inpainting_utils.py."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import ref_harness


def synthetic_png(tmp_path, w=203, h=171):
    g = np.random.RandomState(3)
    small = (g.rand(h // 8 + 1, w // 8 + 1, 3) * 255).astype(np.uint8)
    img = Image.fromarray(small).resize((w, h), Image.BICUBIC)
    p = os.path.join(str(tmp_path), "img.png")
    img.save(p)
    return p


def test_sr_pair_loading_shapes_and_crop(tmp_path):
    from utils.sr_utils import get_baselines, load_LR_HR_imgs_sr, put_in_center
    p = synthetic_png(tmp_path)
    imgs = load_LR_HR_imgs_sr(p, -1, 4, 'CROP')
    assert imgs['orig_np'].shape == (3, 171, 203)
    assert imgs['HR_np'].shape == (3, 160, 192) and imgs['LR_np'].shape == (3, 40, 48)
    # centre crop: offsets (171-160)/2 = 5.5 -> PIL rounds the box; content must come from the original image
    assert imgs['HR_np'].dtype == np.float32 and 0.0 <= imgs['HR_np'].min() and imgs['HR_np'].max() <= 1.0
    bic, sharp, near = get_baselines(imgs['LR_pil'], imgs['HR_pil'])
    assert bic.shape == sharp.shape == near.shape == imgs['HR_np'].shape
    c = put_in_center(imgs['LR_np'], (64, 64))
    assert c.shape == (3, 64, 64) and c.sum() == pytest.approx(imgs['LR_np'].astype(np.float64).sum(), rel=1e-6)
    assert c[:, :12].sum() == 0 and c[:, :, :8].sum() == 0
    no_crop = load_LR_HR_imgs_sr(p, -1, 4, None)
    assert no_crop['HR_np'].shape == (3, 171, 203) and no_crop['LR_np'].shape == (3, 42, 50)


def test_tv_loss_matches_formula():
    from utils.sr_utils import tv_loss
    x = torch.rand(1, 3, 9, 7, dtype=torch.float64)
    want = 0.0
    for c in range(3):
        for i in range(8):
            for j in range(6):
                want += ((x[0, c, i, j + 1] - x[0, c, i, j]) ** 2 + (x[0, c, i + 1, j] - x[0, c, i, j]) ** 2) ** 0.5
    assert float(tv_loss(x)) == pytest.approx(float(want), rel=1e-12)


def test_bernoulli_mask_fraction():
    from utils.inpainting_utils import get_bernoulli_mask
    np.random.seed(0)
    img = Image.fromarray(np.zeros((64, 64, 3), dtype=np.uint8))
    m = np.array(get_bernoulli_mask(img, zero_fraction=0.9))
    assert m.shape == (64, 64, 3) and set(np.unique(m).tolist()) <= {0, 255}   # np_to_pil scales {0, 1} to {0, 255}
    assert 0.05 < (m > 0).mean() < 0.15


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_sr_utils_equal_live_reference():
    from utils import sr_utils as ours
    fname = os.path.join(ref_harness.REF, "data", "sr", "zebra_GT.png")
    mine = ours.load_LR_HR_imgs_sr(fname, -1, 4, 'CROP')
    mine_base = ours.get_baselines(mine['LR_pil'], mine['HR_pil'])
    x = torch.rand(1, 3, 12, 10)
    mine_tv = float(ours.tv_loss(x))
    with ref_harness.reference_modules() as ref:
        import importlib
        rsr = importlib.import_module("utils.sr_utils")
        theirs = rsr.load_LR_HR_imgs_sr(fname, -1, 4, 'CROP')
        theirs_base = rsr.get_baselines(theirs['LR_pil'], theirs['HR_pil'])
        theirs_tv = float(rsr.tv_loss(x))
        theirs_center = rsr.put_in_center(theirs['LR_np'], (128, 160))
    for k in ('orig_np', 'HR_np', 'LR_np'):
        assert np.array_equal(mine[k], theirs[k]), k
    assert mine['HR_np'].shape == (3, 384, 576) and mine['LR_np'].shape == (3, 96, 144)
    for a, b in zip(mine_base, theirs_base):
        assert np.array_equal(a, b)
    assert mine_tv == pytest.approx(theirs_tv, rel=1e-6)
    assert np.array_equal(ours.put_in_center(mine['LR_np'], (128, 160)), theirs_center)


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_inpainting_utils_equal_live_reference():
    from utils import inpainting_utils as ours
    img = Image.open(os.path.join(ref_harness.REF, "data", "inpainting", "kate.png"))
    np.random.seed(4)
    mine = np.array(ours.get_bernoulli_mask(img, 0.8))
    with ref_harness.reference_modules() as ref:
        import importlib
        rin = importlib.import_module("utils.inpainting_utils")
        np.random.seed(4)
        theirs = np.array(rin.get_bernoulli_mask(img, 0.8))
        have_font = os.path.exists('/usr/share/fonts/truetype/freefont/FreeSansBold.ttf')
        if have_font:
            t_theirs = np.array(rin.get_text_mask(img))
    assert np.array_equal(mine, theirs)
    if have_font:
        assert np.array_equal(np.array(ours.get_text_mask(img)), t_theirs)


def test_image_grid_matches_torchvision_make_grid():
    """get_image_grid (used by plot_image_grid, which the notebooks call unconditionally in a few cells) tiles like
    torchvision.utils.make_grid does for the reference (utils/common_utils.py:55-60), without needing torchvision."""
    torchvision = pytest.importorskip("torchvision")
    from utils.common_utils import get_image_grid, plot_image_grid
    rng = np.random.RandomState(0)
    for n, nrow in ((1, 8), (2, 3), (5, 3), (4, 4), (7, 2)):
        ims = [rng.rand(3, 5, 7).astype(np.float32) for _ in range(n)]
        ref = torchvision.utils.make_grid([torch.from_numpy(x) for x in ims], nrow).numpy()
        assert np.array_equal(get_image_grid(ims, nrow), ref)
    mixed = plot_image_grid([rng.rand(3, 4, 4).astype(np.float32), rng.rand(1, 4, 4).astype(np.float32)], 3, 11)
    assert mixed.shape == (3, 8, 14)
