"""models.UNet with the reference's constructor, attribute names and state_dict layout (reference: models/unet.py:32-192).

Builder API only: SURVEY.md section 2 row 15 / section 8f.4 keep UNet outside the accelerated hot path, so this is an
ordinary torch module executed by stock torch ops (on whatever device its tensors live) -- it never touches the
engine, and there is nothing to fall back from.  inpainting.ipynb c14:62-70 builds it with feature_scale=8,
more_layers=1, upsample_mode='deconv', norm_layer=InstanceNorm2d.
"""
import torch
import torch.nn as nn

from .common import conv


class ListModule(nn.Module):
    """Indexable container whose children are named "0", "1", ... (reference: models/unet.py:7-30)."""

    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def __getitem__(self, idx):
        mods = list(self._modules.values())
        if not -len(mods) <= idx < len(mods):
            raise IndexError('index {} is out of range'.format(idx))
        return mods[idx]

    def __iter__(self):
        return iter(self._modules.values())

    def __len__(self):
        return len(self._modules)


def _conv_relu(cin, cout, norm_layer, need_bias, pad):
    mods = [conv(cin, cout, 3, bias=need_bias, pad=pad)]
    if norm_layer is not None:
        mods.append(norm_layer(cout))
    mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class unetConv2(nn.Module):
    """Two (3x3 conv [+ norm] + ReLU) stages: .conv1, .conv2 (reference: models/unet.py:130-151)."""

    def __init__(self, in_size, out_size, norm_layer, need_bias, pad):
        super().__init__()
        self.conv1 = _conv_relu(in_size, out_size, norm_layer, need_bias, pad)
        self.conv2 = _conv_relu(out_size, out_size, norm_layer, need_bias, pad)

    def forward(self, inputs):
        return self.conv2(self.conv1(inputs))


class unetDown(nn.Module):
    """MaxPool 2x2 then unetConv2: .conv, .down (reference: models/unet.py:154-163)."""

    def __init__(self, in_size, out_size, norm_layer, need_bias, pad):
        super().__init__()
        self.conv = unetConv2(in_size, out_size, norm_layer, need_bias, pad)
        self.down = nn.MaxPool2d(2, 2)

    def forward(self, inputs):
        return self.conv(self.down(inputs))


class unetUp(nn.Module):
    """x2 up (ConvTranspose2d 4x4 s2 | Upsample + 3x3 conv), centre-crop the skip tensor, concat, unetConv2 without
    normalisation: .up, .conv (reference: models/unet.py:166-192)."""

    def __init__(self, out_size, upsample_mode, need_bias, pad, same_num_filt=False):
        super().__init__()
        num_filt = out_size if same_num_filt else out_size * 2
        if upsample_mode == 'deconv':
            self.up = nn.ConvTranspose2d(num_filt, out_size, 4, stride=2, padding=1)
        elif upsample_mode in ('bilinear', 'nearest'):
            self.up = nn.Sequential(nn.Upsample(scale_factor=2, mode=upsample_mode),
                                    conv(num_filt, out_size, 3, bias=need_bias, pad=pad))
        else:
            assert False, 'unknown upsample_mode ' + str(upsample_mode)
        self.conv = unetConv2(out_size * 2, out_size, None, need_bias, pad)

    def forward(self, inputs1, inputs2):
        up = self.up(inputs1)
        h, w = up.size(2), up.size(3)
        if (inputs2.size(2), inputs2.size(3)) != (h, w):
            t, l = (inputs2.size(2) - h) // 2, (inputs2.size(3) - w) // 2
            inputs2 = inputs2[:, :, t:t + h, l:l + w]
        return self.conv(torch.cat([up, inputs2], 1))


class UNet(nn.Module):
    """upsample_mode in ['deconv', 'nearest', 'bilinear'], pad in ['zero', 'reflection'] (reference: models/unet.py:32-126)."""

    def __init__(self, num_input_channels=3, num_output_channels=3, feature_scale=4, more_layers=0, concat_x=False,
                 upsample_mode='deconv', pad='zero', norm_layer=nn.InstanceNorm2d, need_sigmoid=True, need_bias=True):
        super().__init__()
        self.feature_scale = feature_scale
        self.more_layers = more_layers
        self.concat_x = concat_x
        f = [c // feature_scale for c in (64, 128, 256, 512, 1024)]
        w = (lambda c: c - num_input_channels) if concat_x else (lambda c: c)   # room for the concatenated input pyramid

        self.start = unetConv2(num_input_channels, w(f[0]), norm_layer, need_bias, pad)
        self.down1 = unetDown(f[0], w(f[1]), norm_layer, need_bias, pad)
        self.down2 = unetDown(f[1], w(f[2]), norm_layer, need_bias, pad)
        self.down3 = unetDown(f[2], w(f[3]), norm_layer, need_bias, pad)
        self.down4 = unetDown(f[3], w(f[4]), norm_layer, need_bias, pad)
        if more_layers > 0:
            self.more_downs = ListModule(*[unetDown(f[4], w(f[4]), norm_layer, need_bias, pad) for _ in range(more_layers)])
            self.more_ups = ListModule(*[unetUp(f[4], upsample_mode, need_bias, pad, same_num_filt=True)
                                         for _ in range(more_layers)])
        self.up4 = unetUp(f[3], upsample_mode, need_bias, pad)
        self.up3 = unetUp(f[2], upsample_mode, need_bias, pad)
        self.up2 = unetUp(f[1], upsample_mode, need_bias, pad)
        self.up1 = unetUp(f[0], upsample_mode, need_bias, pad)
        self.final = conv(f[0], num_output_channels, 1, bias=need_bias, pad=pad)
        if need_sigmoid:
            self.final = nn.Sequential(self.final, nn.Sigmoid())

    def forward(self, inputs):
        pyramid = [inputs]                       # the input at every scale (used when concat_x)
        pool = nn.AvgPool2d(2, 2)
        for _ in range(4 + self.more_layers):
            pyramid.append(pool(pyramid[-1]))

        def with_x(t, k):
            return torch.cat([t, pyramid[k]], 1) if self.concat_x else t

        feats = [with_x(self.start(inputs), 0)]
        for k, stage in enumerate((self.down1, self.down2, self.down3, self.down4), start=1):
            feats.append(with_x(stage(feats[-1]), k))
        x = feats[4]
        if self.more_layers > 0:
            deep = [x]
            for k, stage in enumerate(self.more_downs):
                deep.append(with_x(stage(deep[-1]), k + 5))
            x = self.more_ups[-1](deep[-1], deep[-2])
            # (the reference indexes with an undefined `self.more` here, models/unet.py:116-117: more_layers > 1 raises
            # there; this is the evident intent)
            for idx in range(self.more_layers - 1):
                x = self.more_ups[self.more_layers - idx - 2](x, deep[self.more_layers - idx - 2])
        x = self.up4(x, feats[3])
        x = self.up3(x, feats[2])
        x = self.up2(x, feats[1])
        x = self.up1(x, feats[0])
        return self.final(x)
