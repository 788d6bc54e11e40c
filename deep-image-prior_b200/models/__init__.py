"""`from models import *` surface of the reference (reference: models/__init__.py:1-32): skip, get_net, nn.

The skip network family is the accelerated hot path.  ResNet and UNet keep the reference's
builder API as ordinary torch modules (stock torch ops, never accelerated: SURVEY.md 8f.4); texture_nets (py2-era code
that does not run under py3 in the reference either, models/texture_nets.py:11-13) is not provided.
"""
import torch.nn as nn

from .common import Concat, GenNoise, Swish, act, bn, conv  # noqa: F401
from .downsampler import Downsampler, get_kernel  # noqa: F401
from .skip import SkipNet, allow_torch_execution, skip  # noqa: F401
from .resnet import ResNet  # noqa: F401
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
    if NET_TYPE == 'ResNet':
        # same positional call as the reference (models/__init__.py:9-11, marked TODO there: act_fun receives
        # nn.BatchNorm2d and the construction raises TypeError in the reference too)
        return ResNet(input_depth, 3, 10, 16, 1, nn.BatchNorm2d, False)
    if NET_TYPE == 'UNet':
        return UNet(num_input_channels=input_depth, num_output_channels=3, feature_scale=4, more_layers=0, concat_x=False,
                    upsample_mode=upsample_mode, pad=pad, norm_layer=nn.BatchNorm2d, need_sigmoid=True, need_bias=True)
    if NET_TYPE == 'texture_nets':
        raise NotImplementedError("dip-b200: NET_TYPE='texture_nets' is not provided (the reference's builder does not run "
                                  "under python 3: models/texture_nets.py:11-13 passes float paddings)")
    assert False, 'unknown NET_TYPE ' + str(NET_TYPE)
