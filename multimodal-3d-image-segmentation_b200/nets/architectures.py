"""FNO-family segmentation networks on the B200 kernels (reference: nets/architectures.py:255-429, 511-608).

``NeuralOperatorSeg(..., transform_type='Hartley')`` is HNOSeg: every block runs a Hartley spectral layer WITH its own
transform pair (``HartleyOperator._call3d``, nets/hartley_operator.py:168-271) next to a 1x1x1 conv branch, then SELU
and the concat-skip convolution.  Here a block is four launches of the same kernels as HNOSeg-XS:

    truncated DHT  ->  shared-weight mix + SELU on the retained modes  ->  1x1x1 conv branch  ->
    adjoint DHT whose epilogue adds the conv branch and applies the SELU  ->  concat 1x1x1 conv + bias + SELU

(``selu(0) == 0``, so the SELU the reference applies to the zero-padded spectrum only acts on the retained modes.)
The module tree and ``state_dict`` keys are the reference's (``layers.{i}.op.weight``, ``layers.{i}.conv_branch.weight``,
``layers.{i}.conv_concat.op.*``).  ``transform_type='Fourier'`` (FNOSeg, "FNOSeg3D" of BASELINE.json) swaps the mix for
``FourierOperator.spectral`` (nets/fourier_operator.py: the same truncated transform onto the symmetric mode set, complex
24x24 mixing, no frequency-domain activation); keys ``layers.{i}.op.weight_real / weight_imag``.
"""
from functools import partial
from typing import Union

import numpy as np
import torch
from torch import nn

from .. import ops
from ..plan import get_crop_plan, get_interp_tables, plane_pitch  # noqa: F401
from .fourier_operator import FourierOperator
from .hartley_operator import HartleyOperator
from .nets_utils import ConvNormAct, _is_selu, init_weights_for_snn, spatial_padcrop


class NeuralOperatorBlock(nn.Module):
    """FNO / HNO block (reference :551-608 with _TransBlock.forward :521-548)."""

    def __init__(self, in_channels, out_channels, num_modes, transform_type, weights_type='shared', ndim=5,
                 activation='selu', device=None, use_conv_branch=True, use_bias_conv_branch=False, use_block_skip=True,
                 use_block_concat=True):
        super().__init__()
        assert transform_type in ('Fourier', 'Hartley')
        if ndim != 5:
            raise NotImplementedError('hno_b200 NeuralOperatorBlock supports 3-D (ndim=5) only')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 implements the SELU (self-normalising) variant only')
        if not use_conv_branch:
            raise NotImplementedError('hno_b200: NeuralOperatorBlock without the conv branch is not supported')
        self.use_block_skip = use_block_skip
        op = FourierOperator if transform_type == 'Fourier' else HartleyOperator
        self.op = op(in_channels, out_channels, num_modes, use_bias=False, weights_type=weights_type, ndim=ndim,
                     device=device)
        self.transform_type = transform_type
        self.conv_branch = nn.Conv3d(in_channels, out_channels, kernel_size=1, bias=use_bias_conv_branch, device=device)
        self.normalization = None
        self.activation = nn.functional.selu
        self.conv_concat = None
        if use_block_skip and use_block_concat:
            self.conv_concat = ConvNormAct(in_channels + out_channels, out_channels, use_bias=True,
                                           activation=activation, ndim=ndim, device=device)

    def forward(self, x):
        if x.is_meta:
            y = self.activation(self.op(x) + self.conv_branch(x))
            if self.use_block_skip:
                y = self.conv_concat(torch.cat([y, x], 1)) if self.conv_concat is not None else y + x
            return y
        x = x.contiguous()
        spatial = tuple(x.shape[2:])
        wb = self.conv_branch.weight
        t = ops.PointwiseConv.apply(x, None, wb.view(wb.shape[0], -1), self.conv_branch.bias, 0, False)
        # Fourier: no activation in the frequency domain (fourier_operator.py:148-211); Hartley: mix + SELU on the modes
        z, plan = self.op.spectral(x)
        y = ops.AddIDHTSelu.apply(t, z, plan)
        if self.use_block_skip:
            if self.conv_concat is not None:
                op = self.conv_concat.op
                return ops.PointwiseConv.apply(y, x, op.weight.view(op.weight.shape[0], -1), op.bias, 1, False)
            return y + x
        return y


class NeuralOperatorSeg(nn.Module):
    """FNO / FNOSeg / HNOSeg family (reference :356-429 over _TransSeg :255-353), shared weights."""

    def __init__(self, in_channels, out_channels, filters, num_transform_blocks, num_modes, transform_type,
                 weights_type='shared', use_resize=True, use_deep_supervision=False, use_bias_conv_branch=False,
                 use_block_skip=True, use_block_concat=True, activation='selu',
                 output_activation: Union[str, callable] = 'softmax', ndim=5, device=None):
        super().__init__()
        assert transform_type in ('Fourier', 'Hartley')
        assert ndim in (4, 5)
        if ndim != 5:
            raise NotImplementedError('hno_b200 NeuralOperatorSeg supports 3-D (ndim=5) only')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 implements the SELU (self-normalising) variant only')
        if use_deep_supervision:
            raise NotImplementedError('hno_b200: deep supervision is not supported')
        if not use_resize:
            raise NotImplementedError('hno_b200: use_resize=False is not supported')
        if output_activation not in ('softmax', None):
            raise NotImplementedError("hno_b200: output_activation must be 'softmax' or None")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.filters = filters
        self.num_transform_blocks = num_transform_blocks
        self.num_modes = (int(num_modes),) * 3 if np.isscalar(num_modes) else tuple(int(m) for m in num_modes)
        self.transform_type = transform_type
        self.weights_type = weights_type
        self.use_resize = use_resize
        self.use_deep_supervision = use_deep_supervision
        self.use_bias_conv_branch = use_bias_conv_branch
        self.use_block_skip = use_block_skip
        self.use_block_concat = use_block_concat
        self.activation = activation
        self.output_activation = output_activation
        self.ndim = ndim
        self.device = device
        self.block = partial(NeuralOperatorBlock, num_modes=self.num_modes, transform_type=transform_type,
                             weights_type=weights_type, ndim=ndim, activation=activation, device=device,
                             use_bias_conv_branch=use_bias_conv_branch, use_block_skip=use_block_skip,
                             use_block_concat=use_block_concat)
        self.conv_in = ConvNormAct(in_channels, filters, kernel_size=2, stride=2, use_bias=True, activation=activation,
                                   ndim=ndim, device=device)
        self.conv1 = ConvNormAct(filters, filters, use_bias=True, activation=activation, ndim=ndim, device=device)
        self.layers = nn.ModuleList([self.block(filters, filters) for _ in range(num_transform_blocks)])
        self.conv_ds = None
        self.conv_out = nn.Conv3d(filters, out_channels, kernel_size=1, bias=False, device=device)
        self.apply(init_weights_for_snn)

    def forward(self, x):
        image_size = tuple(x.shape[2:])
        if x.is_meta:
            y = self.conv1(self.conv_in(x))
            for layer in self.layers:
                y = layer(y)
            y = self.conv_out(nn.functional.interpolate(y, size=image_size, mode='trilinear'))
            return torch.softmax(y, 1) if self.output_activation == 'softmax' else y
        x = self.conv1(self.conv_in(x))
        for layer in self.layers:
            x = layer(x)
        # conv_out has no bias and trilinear weights sum to one: the 1x1x1 conv commutes with the interpolation, so
        # it runs at low resolution and only `out_channels` channels are up-sampled (head_kernels.cu)
        w = self.conv_out.weight
        ll = ops.PointwiseConv.apply(x, None, w.view(w.shape[0], -1), None, 0, False)
        tables = get_interp_tables(tuple(x.shape[2:]), image_size, x.device)
        y = ops.HeadUpsample.apply(ll, tables, 1 if self.output_activation == 'softmax' else 0)
        return spatial_padcrop(y, image_size)
