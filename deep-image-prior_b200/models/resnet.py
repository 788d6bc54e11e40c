"""models.ResNet with the reference's constructor and state_dict layout (reference: models/resnet.py:9-96).

Builder API only (outside the accelerated hot path): an ordinary torch module run
by stock torch ops; inpainting.ipynb c14:72-77 builds ResNet(input_depth, 3, 8, 32, need_sigmoid=True,
act_fun='LeakyReLU').
"""
import torch.nn as nn

from .common import act, conv


class ResidualSequential(nn.Sequential):
    """out = body(x) + x, x centre-cropped to the body's output size (reference: models/resnet.py:9-24; the crop there
    uses float slice bounds, a py2 leftover that raises under py3 -- integer bounds here)."""

    def forward(self, x):
        out = super().forward(x)
        if out.size(2) != x.size(2) or out.size(3) != x.size(3):
            t, l = (x.size(2) - out.size(2)) // 2, (x.size(3) - out.size(3)) // 2
            x = x[:, :, t:t + out.size(2), l:l + out.size(3)]
        return out + x


def get_block(num_channels, norm_layer, act_fun):
    """conv3x3 - norm - act - conv3x3 - norm, zero padding, no conv bias (reference: models/resnet.py:33-41)."""
    return [nn.Conv2d(num_channels, num_channels, 3, 1, 1, bias=False), norm_layer(num_channels, affine=True), act(act_fun),
            nn.Conv2d(num_channels, num_channels, 3, 1, 1, bias=False), norm_layer(num_channels, affine=True)]


class ResNet(nn.Module):
    def __init__(self, num_input_channels, num_output_channels, num_blocks, num_channels, need_residual=True,
                 act_fun='LeakyReLU', need_sigmoid=True, norm_layer=nn.BatchNorm2d, pad='reflection'):
        super().__init__()
        block_type = ResidualSequential if need_residual else nn.Sequential
        layers = [conv(num_input_channels, num_channels, 3, stride=1, bias=True, pad=pad), act(act_fun)]
        layers += [block_type(*get_block(num_channels, norm_layer, act_fun)) for _ in range(num_blocks)]
        layers += [nn.Conv2d(num_channels, num_channels, 3, 1, 1), norm_layer(num_channels, affine=True)]
        # the reference appends the sigmoid unconditionally (need_sigmoid is accepted and ignored, models/resnet.py:84-87)
        layers += [conv(num_channels, num_output_channels, 3, 1, bias=True, pad=pad), nn.Sigmoid()]
        self.model = nn.Sequential(*layers)

    def forward(self, input):
        return self.model(input)

    def eval(self):
        self.model.eval()
