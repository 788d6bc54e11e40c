"""Task utilities with the reference's names (reference: utils/common_utils.py).

`from utils.common_utils import *` also re-exports torch, nn, np, Image, PIL, torchvision and plt like the reference
does (notebooks rely on it); matplotlib is imported lazily because it is optional here.
"""
import sys

import numpy as np
import PIL
import torch
import torch.nn as nn
from PIL import Image

try:
    import torchvision
except Exception:  # pragma: no cover
    torchvision = None
try:
    import matplotlib.pyplot as plt
except Exception:  # matplotlib is not installed on the build/GPU image; plotting helpers then no-op
    plt = None

if not hasattr(Image, 'ANTIALIAS'):  # removed in Pillow >= 10; same filter
    Image.ANTIALIAS = Image.LANCZOS


def crop_image(img, d=32):
    """Centre-crops a PIL image so both sides are divisible by `d` (reference: utils/common_utils.py:13-27)."""
    w, h = img.size[0] - img.size[0] % d, img.size[1] - img.size[1] % d
    box = [int((img.size[0] - w) / 2), int((img.size[1] - h) / 2), int((img.size[0] + w) / 2),
           int((img.size[1] + h) / 2)]
    return img.crop(box)


def get_params(opt_over, net, net_input, downsampler=None):
    """Parameters to optimise over: comma separated subset of net,down,input (reference: :29-53)."""
    params = []
    for opt in opt_over.split(','):
        if opt == 'net':
            params += [x for x in net.parameters()]
        elif opt == 'down':
            assert downsampler is not None
            params = [x for x in downsampler.parameters()]
        elif opt == 'input':
            net_input.requires_grad = True
            params += [net_input]
        else:
            assert False, 'what is it?'
    return params


def get_image_grid(images_np, nrow=8):
    """C x H x W arrays tiled into one C x H' x W' array, `nrow` images per row, 2-pixel black gutters (what the reference
    gets from torchvision.utils.make_grid, utils/common_utils.py:55-60; a single image comes back as it is)."""
    imgs = [np.asarray(x) for x in images_np]
    if len(imgs) == 1:
        return imgs[0]
    gap = 2
    c, h, w = imgs[0].shape
    cols = min(nrow, len(imgs))
    rows = -(-len(imgs) // cols)
    canvas = np.zeros((c, rows * (h + gap) + gap, cols * (w + gap) + gap), dtype=imgs[0].dtype)
    for k, im in enumerate(imgs):
        top, left = (k // cols) * (h + gap) + gap, (k % cols) * (w + gap) + gap
        canvas[:, top:top + h, left:left + w] = im
    return canvas


def plot_image_grid(images_np, nrow=8, factor=1, interpolation='lanczos'):
    """Display helper of the notebooks (reference: utils/common_utils.py:62-87): grey images are promoted to 3 channels when
    mixed with colour ones, the tiled grid is shown with matplotlib when it is installed, and returned either way."""
    depth = max(im.shape[0] for im in images_np)
    assert depth in (1, 3), "images should have 1 or 3 channels"
    tiles = [im if im.shape[0] == depth else np.repeat(im, 3, axis=0) for im in images_np]
    grid = get_image_grid(tiles, nrow)
    if plt is not None:
        plt.figure(figsize=(len(tiles) + factor, 12 + factor))
        shown = grid[0] if depth == 1 else np.moveaxis(grid, 0, -1)
        plt.imshow(shown, interpolation=interpolation, **({'cmap': 'gray'} if depth == 1 else {}))
        plt.show()
    return grid


def load(path):
    return Image.open(path)


def get_image(path, imsize=-1):
    """Loads an image, optionally resized; returns (PIL image, C x H x W float array) (reference: :94-114)."""
    img = load(path)
    if isinstance(imsize, int):
        imsize = (imsize, imsize)
    if imsize[0] != -1 and img.size != imsize:
        img = img.resize(imsize, Image.BICUBIC if imsize[0] > img.size[0] else Image.ANTIALIAS)
    return img, pil_to_np(img)


def fill_noise(x, noise_type):
    if noise_type == 'u':
        x.uniform_()
    elif noise_type == 'n':
        x.normal_()
    else:
        assert False


def get_noise(input_depth, method, spatial_size, noise_type='u', var=1. / 10):
    """1 x input_depth x H x W input tensor: scaled noise or a meshgrid (reference: :127-153)."""
    if isinstance(spatial_size, int):
        spatial_size = (spatial_size, spatial_size)
    if method == 'noise':
        net_input = torch.zeros([1, input_depth, spatial_size[0], spatial_size[1]])
        fill_noise(net_input, noise_type)
        net_input *= var
    elif method == 'meshgrid':
        assert input_depth == 2
        X, Y = np.meshgrid(np.arange(0, spatial_size[1]) / float(spatial_size[1] - 1),
                           np.arange(0, spatial_size[0]) / float(spatial_size[0] - 1))
        net_input = np_to_torch(np.concatenate([X[None, :], Y[None, :]]))
    else:
        assert False
    return net_input


def pil_to_np(img_PIL):
    ar = np.array(img_PIL)
    ar = ar.transpose(2, 0, 1) if ar.ndim == 3 else ar[None, ...]
    return ar.astype(np.float32) / 255.


def np_to_pil(img_np):
    ar = np.clip(img_np * 255, 0, 255).astype(np.uint8)
    ar = ar[0] if img_np.shape[0] == 1 else ar.transpose(1, 2, 0)
    return Image.fromarray(ar)


def np_to_torch(img_np):
    return torch.from_numpy(img_np)[None, :]


def torch_to_np(img_var):
    return img_var.detach().cpu().numpy()[0]


def optimize(optimizer_type, parameters, closure, LR, num_iter):
    """The optimisation loop (reference: utils/common_utils.py:198-232).

    'adam': zero_grad(); closure(); step() `num_iter` times.  For CUDA parameters the step is the engine's fused
    multi-tensor Adam (dip_adam_step), arithmetic identical to torch.optim.Adam; the closure stays an opaque callable.
    """
    parameters = list(parameters)
    on_gpu = len(parameters) > 0 and all(p.is_cuda and p.dtype == torch.float32 for p in parameters)

    def make_adam(lr):
        if on_gpu:
            import dip_engine
            return dip_engine.FusedAdam(parameters, lr=lr)
        from models.skip import _ALLOW_TORCH
        if not _ALLOW_TORCH:
            raise RuntimeError("dip-b200: optimize() runs on CUDA float32 parameters; there is no CPU fallback "
                               "(models.allow_torch_execution(True) opts in to stock torch)")
        return torch.optim.Adam(parameters, lr=lr)

    if optimizer_type == 'LBFGS':
        optimizer = make_adam(0.001)
        for j in range(100):
            optimizer.zero_grad()
            closure()
            optimizer.step()
        print('Starting optimization with LBFGS')

        def closure2():
            optimizer.zero_grad()
            return closure()
        optimizer = torch.optim.LBFGS(parameters, max_iter=num_iter, lr=LR, tolerance_grad=-1, tolerance_change=-1)
        optimizer.step(closure2)
    elif optimizer_type == 'adam':
        print('Starting optimization with ADAM')
        optimizer = make_adam(LR)
        for j in range(num_iter):
            optimizer.zero_grad()
            closure()
            optimizer.step()
    else:
        assert False
