"""reference: utils/denoising_utils.py"""
import os  # noqa: F401

from .common_utils import *  # noqa: F401,F403
from .common_utils import np, np_to_pil


def get_noisy_image(img_np, sigma):
    """clip(img + N(0, sigma), 0, 1) as float32, numpy global RNG (reference: utils/denoising_utils.py:6-16)."""
    img_noisy_np = np.clip(img_np + np.random.normal(scale=sigma, size=img_np.shape), 0, 1).astype(np.float32)
    return np_to_pil(img_noisy_np), img_noisy_np
