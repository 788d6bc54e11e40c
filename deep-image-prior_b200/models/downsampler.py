"""Anti-aliased strided downsampling operator (reference: models/downsampler.py:5-135).

Same constructor and numerics as the reference: a fixed K x K Lanczos / Gauss / box filter applied per channel with
stride `factor`, optional replication padding.  The filter is separable (outer product of 1-D taps), which is how it
is built here.  CUDA tensors run on the engine's stencil kernels (dip_lanczos_down_fwd / _bwd in include/dip.h, one
autograd node); there is no CPU path unless models.allow_torch_execution(True) opts in to the stock torch modules.
"""
import numpy as np
import torch
import torch.nn as nn

_PRESETS = {
    # name: (kernel_type, support, width(factor), sigma)
    'lanczos2': ('lanczos', 2, lambda f: 4 * f + 1, None),
    'lanczos3': ('lanczos', 3, lambda f: 6 * f + 1, None),
    'gauss12': ('gauss', None, lambda f: 7, 0.5),
    'gauss1sq2': ('gauss', None, lambda f: 9, 1. / np.sqrt(2)),
}


def _lanczos_taps(n, center, factor, support, half_phase):
    i = np.arange(1, n + 1, dtype=np.float64)
    d = np.abs(i + (0.5 if half_phase else 0.0) - center) / factor
    out = np.ones(n, dtype=np.float64)
    nz = d != 0
    dn = d[nz]
    out[nz] = support * np.sin(np.pi * dn) * np.sin(np.pi * dn / support) / (np.pi * np.pi * dn * dn)
    return out


def get_kernel(factor, kernel_type, phase, kernel_width, support=None, sigma=None):
    """K x K float64 filter, normalised to sum 1 (reference: models/downsampler.py:73-135)."""
    assert kernel_type in ['lanczos', 'gauss', 'box']
    n = kernel_width - 1 if (phase == 0.5 and kernel_type != 'box') else kernel_width
    if kernel_type == 'box':
        assert phase == 0.5, 'Box filter is always half-phased'
        kernel = np.full([n, n], 1. / (kernel_width * kernel_width))
    elif kernel_type == 'gauss':
        assert sigma, 'sigma is not specified'
        assert phase != 0.5, 'phase 1/2 for gauss not implemented'
        center = (kernel_width + 1.) / 2.
        d = (np.arange(1, n + 1, dtype=np.float64) - center) / 2.
        g = np.exp(-(d * d) / (2 * sigma * sigma))
        kernel = np.outer(g, g) / (2. * np.pi * sigma * sigma)
    else:
        assert support, 'support is not specified'
        center = (kernel_width + 1) / 2.
        t = _lanczos_taps(n, center, factor, support, phase == 0.5)
        kernel = np.outer(t, t)
    kernel /= kernel.sum()
    return kernel


class _DownFn(torch.autograd.Function):
    """out_LR = downsampler(out_HR) on the engine (reference: models/downsampler.py:58-71)."""

    @staticmethod
    def forward(ctx, x, kern, factor, pad):
        import dip_engine as de
        ctx.kern, ctx.factor, ctx.pad, ctx.hw = kern, factor, pad, (int(x.shape[2]), int(x.shape[3]))
        return de.lanczos_down_fwd(x, kern, factor, pad)

    @staticmethod
    def backward(ctx, dy):
        import dip_engine as de
        return de.lanczos_down_bwd(dy, ctx.kern, ctx.factor, ctx.pad, *ctx.hw), None, None, None


class Downsampler(nn.Module):
    def __init__(self, n_planes, factor, kernel_type, phase=0, kernel_width=None, support=None, sigma=None,
                 preserve_size=False):
        super().__init__()
        assert phase in [0, 0.5], 'phase should be 0 or 0.5'
        if kernel_type in _PRESETS:
            kernel_type_, support_, width_fn, sigma_ = _PRESETS[kernel_type]
            support = support_ if support_ is not None else support
            sigma = sigma_ if sigma_ is not None else sigma
            kernel_width = width_fn(factor)
        elif kernel_type in ['lanczos', 'gauss', 'box']:
            kernel_type_ = kernel_type
        else:
            assert False, 'wrong name kernel'
        self.kernel = get_kernel(factor, kernel_type_, phase, kernel_width, support=support, sigma=sigma)
        op = nn.Conv2d(n_planes, n_planes, kernel_size=self.kernel.shape, stride=factor, padding=0)
        with torch.no_grad():
            op.weight.zero_()
            op.bias.zero_()
            k = torch.from_numpy(self.kernel)
            for c in range(n_planes):
                op.weight[c, c] = k
        self.downsampler_ = op
        self.factor = factor
        self.pad = 0
        if preserve_size:
            ks = self.kernel.shape[0]
            pad = int((ks - 1) / 2.) if ks % 2 == 1 else int((ks - factor) / 2.)
            self.padding = nn.ReplicationPad2d(pad)
            self.pad = pad
        self.preserve_size = preserve_size

    def forward(self, input):
        if input.is_cuda:
            # engine path: the K x K taps of plane 0 (the weight is plane-diagonal by construction; the conv bias is
            # zero by construction and not applied -- optimising the operator itself, OPT_OVER='down', is out of scope)
            if input.dtype != torch.float32 or self.downsampler_.weight.device != input.device:
                raise RuntimeError("dip-b200: Downsampler needs float32 CUDA tensors on the module's device "
                                   "(downsampler.type(torch.cuda.FloatTensor))")
            if input.dim() != 4 or input.shape[1] != self.downsampler_.weight.shape[0]:
                raise ValueError("dip-b200: Downsampler built for %d planes got an input of shape %s"
                                 % (self.downsampler_.weight.shape[0], tuple(input.shape)))
            kern = self.downsampler_.weight.detach()[0, 0]
            return _DownFn.apply(input, kern, self.factor, self.pad)
        from .skip import _ALLOW_TORCH
        if not _ALLOW_TORCH:
            raise RuntimeError("dip-b200: Downsampler runs on CUDA tensors; there is no CPU fallback "
                               "(models.allow_torch_execution(True) opts in to stock torch)")
        x = self.padding(input) if self.preserve_size else input
        self.x = x
        return self.downsampler_(x)
