"""Super-resolution task utilities with the reference's names (reference: utils/sr_utils.py:1-94); host-side data
preparation around the hot path (super-resolution.ipynb c4-c6: image pair loading; c10:10: optional TV term)."""
from .common_utils import *  # noqa: F401,F403  (the notebooks rely on the star re-export, like the reference)
import PIL.ImageFilter


def put_in_center(img_np, target_size):
    """Pastes a C x h x w array into the middle of a zero 3 x H x W canvas (reference: utils/sr_utils.py:3-15)."""
    canvas = np.zeros([3, target_size[0], target_size[1]])
    h, w = img_np.shape[1], img_np.shape[2]
    top, left = int((target_size[0] - h) / 2), int((target_size[1] - w) / 2)
    bottom, right = int((target_size[0] + h) / 2), int((target_size[1] + w) / 2)
    canvas[:, top:bottom, left:right] = img_np
    return canvas


def load_LR_HR_imgs_sr(fname, imsize, factor, enforse_div32=None):
    """Loads an image, optionally resizes it, centre-crops it to multiples of 32 ('CROP') and makes the low-resolution
    partner with PIL's antialiased resize (reference: utils/sr_utils.py:18-68).  Returns the reference's dict:
    orig_pil/orig_np, HR_pil/HR_np, LR_pil/LR_np."""
    orig_pil, orig_np = get_image(fname, -1)
    if imsize != -1:
        orig_pil, orig_np = get_image(fname, imsize)
    hr_pil, hr_np = orig_pil, orig_np
    if enforse_div32 == 'CROP':
        W0, H0 = orig_pil.size
        Wc, Hc = W0 - W0 % 32, H0 - H0 % 32
        hr_pil = orig_pil.crop([(W0 - Wc) / 2, (H0 - Hc) / 2, (W0 + Wc) / 2, (H0 + Hc) / 2])
        hr_np = pil_to_np(hr_pil)
    lr_pil = hr_pil.resize([hr_pil.size[0] // factor, hr_pil.size[1] // factor], Image.ANTIALIAS)
    print('HR and LR resolutions: %s, %s' % (str(hr_pil.size), str(lr_pil.size)))
    return {'orig_pil': orig_pil, 'orig_np': orig_np, 'LR_pil': lr_pil, 'LR_np': pil_to_np(lr_pil),
            'HR_pil': hr_pil, 'HR_np': hr_np}


def get_baselines(img_LR_pil, img_HR_pil):
    """Bicubic, unsharp-masked bicubic and nearest-neighbour upscalings of the LR image (reference: :71-82)."""
    size = img_HR_pil.size
    bicubic = img_LR_pil.resize(size, Image.BICUBIC)
    nearest = img_LR_pil.resize(size, Image.NEAREST)
    sharp = bicubic.filter(PIL.ImageFilter.UnsharpMask())
    return pil_to_np(bicubic), pil_to_np(sharp), pil_to_np(nearest)


def tv_loss(x, beta=0.5):
    """Total-variation prior sum((dh^2 + dw^2)^beta) over the common (H-1) x (W-1) window (reference: :86-94)."""
    dh = (x[:, :, :, 1:] - x[:, :, :, :-1]) ** 2
    dw = (x[:, :, 1:, :] - x[:, :, :-1, :]) ** 2
    return torch.sum((dh[:, :, :-1] + dw[:, :, :, :-1]) ** beta)
