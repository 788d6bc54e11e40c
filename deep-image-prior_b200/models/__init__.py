"""`from models import *` surface of the reference (reference: models/__init__.py:1-32): skip, get_net, nn.

Round-1 scope: the skip network family is the hot path.  The other builders of the reference
(ResNet, UNet, texture_nets) are outside the accelerated path and not provided yet; asking for them raises.
"""
import torch.nn as nn

from .common import Concat, GenNoise, Swish, act, bn, conv  # noqa: F401
from .downsampler import Downsampler, get_kernel  # noqa: F401
from .skip import SkipNet, allow_torch_execution, skip  # noqa: F401
from .resnet import ResNet  # noqa: F401  (import shims: the notebooks import these names; building them raises)
from .unet import UNet  # noqa: F401


def get_net(input_depth, NET_TYPE, pad, upsample_mode, n_channels=3, act_fun='LeakyReLU', skip_n33d=128, skip_n33u=128,
            skip_n11=4, num_scales=5, downsample_mode='stride'):
    """Network factory with the reference's signature (reference: models/__init__.py:8-32)."""
    if NET_TYPE == 'skip':
        as_list = lambda v: [v] * num_scales if isinstance(v, int) else v  # noqa: E731
        return skip(input_depth, n_channels, num_channels_down=as_list(skip_n33d), num_channels_up=as_list(skip_n33u),
                    num_channels_skip=as_list(skip_n11), upsample_mode=upsample_mode, downsample_mode=downsample_mode,
                    need_sigmoid=True, need_bias=True, pad=pad, act_fun=act_fun)
    if NET_TYPE == 'identity':
        assert input_depth == 3
        return nn.Sequential()
    if NET_TYPE in ('ResNet', 'UNet', 'texture_nets'):
        raise NotImplementedError("dip-b200: NET_TYPE=%r is outside the accelerated hot path (SURVEY.md section 8f) and "
                                  "is not provided in this round" % NET_TYPE)
    assert False, 'unknown NET_TYPE ' + str(NET_TYPE)
