"""Inpainting task utilities with the reference's names (reference: utils/inpainting_utils.py:1-22): synthetic masks
for the inpainting notebook (the kate / vase / library masks themselves ship as images)."""
import PIL.ImageDraw as ImageDraw
import PIL.ImageFont as ImageFont

from .common_utils import *  # noqa: F401,F403

_FONT = '/usr/share/fonts/truetype/freefont/FreeSansBold.ttf'


def get_text_mask(for_image, sz=20):
    """White mask image of for_image's size with the words "hello world" in black at (128, 128) (reference: :7-16)."""
    font = ImageFont.truetype(_FONT, sz)
    mask = Image.fromarray(np.array(for_image) * 0 + 255)
    ImageDraw.Draw(mask).text((128, 128), "hello world", font=font, fill='rgb(0, 0, 0)')
    return mask


def get_bernoulli_mask(for_image, zero_fraction=0.95):
    """Mask that keeps each element with probability 1 - zero_fraction (reference: :18-22)."""
    keep = np.random.random_sample(size=pil_to_np(for_image).shape) > zero_fraction
    return np_to_pil(keep.astype(int))
