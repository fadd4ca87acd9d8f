"""Building blocks shared by the architectures (reference: nets/nets_utils.py)."""
import numpy as np
import torch
from torch import nn

from .. import ops


def get_spatial_padcrop(x, target_shape):
    """Per-axis (pad, crop) amounts, PyTorch order (last axis first); the odd voxel goes to the upper side."""
    shape = tuple(x.shape[2:])
    nd = len(shape)
    padding, cropping = [0] * (2 * nd), [0] * (2 * nd)
    for pos, (t, s) in enumerate(zip(reversed(tuple(target_shape)), reversed(shape))):
        diff = t - s
        lo = abs(diff) // 2
        hi = abs(diff) - lo
        if diff > 0:
            padding[2 * pos], padding[2 * pos + 1] = lo, hi
        elif diff < 0:
            cropping[2 * pos], cropping[2 * pos + 1] = lo, hi
    return padding, cropping


def spatial_padcrop(x, target_shape):
    """Centre pad and/or crop to `target_shape` (reference :22-57).  A no-op when the shapes agree."""
    assert x.ndim in (3, 4, 5) and x.ndim == len(target_shape) + 2
    padding, cropping = get_spatial_padcrop(x, target_shape)
    if any(padding):
        x = nn.functional.pad(x, padding)
    if any(cropping):
        for pos in range(x.ndim - 2):
            lo, hi = cropping[2 * pos], cropping[2 * pos + 1]
            axis = x.ndim - 1 - pos
            x = x.narrow(axis, lo, x.shape[axis] - lo - hi)
    return x


def init_weights_for_snn(module):
    """SNN initialisation (reference :102-117): kaiming-normal 'linear' weights, bias ~ U(-1e-3, 1e-3)."""
    from .fourier_operator import FourierOperator
    from .hartley_operator import HartleyOperator
    if isinstance(module, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d, HartleyOperator)):
        nn.init.kaiming_normal_(module.weight, nonlinearity='linear')
        if module.bias is not None:
            nn.init.uniform_(module.bias, -0.001, 0.001)
    elif isinstance(module, FourierOperator):
        nn.init.kaiming_normal_(module.weight_real, nonlinearity='linear')
        nn.init.kaiming_normal_(module.weight_imag, nonlinearity='linear')


def _is_selu(activation):
    return activation == 'selu' or activation is nn.functional.selu


class ConvNormAct(nn.Module):
    """Conv3d + SELU on the CUDA kernels: kernel 1 / stride 1 (pointwise) or kernel 2 / stride 2 (the stem).

    ``self.op`` is a real ``nn.Conv3d`` used purely as the parameter holder, so initialisation draws and
    ``state_dict`` keys (``op.weight``, ``op.bias``) are those of the reference (nets_utils.py:136-174).
    """

    def __init__(self, in_channels, out_channels, *, kernel_size=1, stride=1, use_bias=True, activation='selu',
                 use_snn=True, ndim=5, device=None):
        super().__init__()
        if ndim != 5:
            raise NotImplementedError('hno_b200 ConvNormAct supports 3-D (ndim=5) only')
        if not (use_snn and _is_selu(activation)):
            raise NotImplementedError('hno_b200 implements the self-normalising (SELU, no GroupNorm) variant only')
        if np.isscalar(kernel_size) and np.isscalar(stride) and (kernel_size, stride) in ((1, 1), (2, 2)):
            pass
        else:
            raise NotImplementedError('hno_b200 ConvNormAct supports kernel_size=1/stride=1 and kernel_size=2/stride=2')
        padding = 'same' if stride == 1 else kernel_size // 2
        self.op = nn.Conv3d(in_channels, out_channels, kernel_size, stride, padding, bias=use_bias, device=device)
        self.normalization = None
        self.activation = nn.functional.selu
        self.kernel_size = kernel_size

    def forward(self, x):
        if x.is_meta:
            return self.activation(self.op(x))
        if self.kernel_size == 2:
            return ops.StemConv.apply(x, self.op.weight, self.op.bias)
        return ops.PointwiseConv.apply(x, None, self.op.weight, self.op.bias, 1, False)
