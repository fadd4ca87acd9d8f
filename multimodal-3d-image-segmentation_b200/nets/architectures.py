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


class HartleyMHABlock(nn.Module):
    """HartleyMHA block (reference :611-635 with _TransBlock.forward :521-548): Hartley multi-head attention next to a
    1x1x1 conv branch, SELU, concat-skip convolution."""

    def __init__(self, in_channels, key_dim, num_heads, num_modes, patch_size, attention_activation, ndim, activation,
                 device, use_conv_branch=True, use_bias_conv_branch=False, use_block_skip=True, use_block_concat=True):
        super().__init__()
        from .hartley_mha import HartleyMultiHeadAttention
        if ndim != 5:
            raise NotImplementedError('hno_b200 HartleyMHABlock supports 3-D (ndim=5) only')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 implements the SELU (self-normalising) variant only')
        if not use_conv_branch:
            raise NotImplementedError('hno_b200: HartleyMHABlock without the conv branch is not supported')
        self.use_block_skip = use_block_skip
        self.op = HartleyMultiHeadAttention(in_channels, key_dim, num_heads, num_modes, patch_size, attention_activation,
                                            ndim=ndim, device=device)
        self.conv_branch = nn.Conv3d(in_channels, key_dim, kernel_size=1, bias=use_bias_conv_branch, device=device)
        self.normalization = None
        self.activation = nn.functional.selu
        self.conv_concat = None
        if use_block_skip and use_block_concat:
            self.conv_concat = ConvNormAct(in_channels + key_dim, key_dim, use_bias=True, activation=activation,
                                           ndim=ndim, device=device)

    def forward(self, x):
        if x.is_meta:
            y = self.activation(self.op(x) + self.conv_branch(x))
            if self.use_block_skip:
                y = self.conv_concat(torch.cat([y, x], 1)) if self.conv_concat is not None else y + x
            return y
        x = x.contiguous()
        wb = self.conv_branch.weight
        t = ops.PointwiseConv.apply(x, None, wb.view(wb.shape[0], -1), self.conv_branch.bias, 0, False)
        z, plan = self.op.spectral(x)
        y = ops.AddIDHTSelu.apply(t, z, plan)
        if self.use_block_skip:
            if self.conv_concat is not None:
                op = self.conv_concat.op
                return ops.PointwiseConv.apply(y, x, op.weight.view(op.weight.shape[0], -1), op.bias, 1, False)
            return y + x
        return y


class _TransSeg(nn.Module):
    """Common body of NeuralOperatorSeg and HartleyMHASeg (reference :255-353): stem, conv1, the transform blocks, optional
    deep supervision (conv_ds over the concatenation of conv1's and every block's output), up-sampling, conv_out, softmax.

    forward() runs the whole network through engine.TransSegEngine (one autograd node, planar activations, in-place
    skip / deep-supervision gradients) when every block is of a kind the engine knows; otherwise block by block through the
    stand-alone modules (forward_modular, also kept as an independent cross-check)."""

    def _init_common(self, in_channels, out_channels, filters, num_transform_blocks, use_resize, use_deep_supervision,
                     activation, output_activation, ndim, device):
        assert ndim in (4, 5)
        if ndim != 5:
            raise NotImplementedError(f'hno_b200 {type(self).__name__} supports 3-D (ndim=5) only')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 implements the SELU (self-normalising) variant only')
        if output_activation not in ('softmax', None):
            raise NotImplementedError("hno_b200: output_activation must be 'softmax' or None")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.filters = filters
        self.num_transform_blocks = num_transform_blocks
        self.use_resize = use_resize
        self.use_deep_supervision = use_deep_supervision
        self.activation = activation
        self.output_activation = output_activation
        self.output_activation_name = output_activation
        self.ndim = ndim
        self.device = device
        self._engine = None

    def create_layers(self):
        f = self.filters
        self.conv_in = None
        cur = self.in_channels
        if self.use_resize:  # reference :286-289; without it the blocks run at the image resolution
            self.conv_in = ConvNormAct(cur, f, kernel_size=2, stride=2, use_bias=True, activation=self.activation,
                                       ndim=self.ndim, device=self.device)
            cur = f
        self.conv1 = ConvNormAct(cur, f, use_bias=True, activation=self.activation, ndim=self.ndim, device=self.device)
        self.layers = nn.ModuleList([self.block(f, f) for _ in range(self.num_transform_blocks)])
        self.conv_ds = None
        cur = f
        if self.use_deep_supervision:
            # reference :306-311: "to avoid OOM" the concatenation is reduced to out_channels before the up-sampling
            self.conv_ds = ConvNormAct(f * (self.num_transform_blocks + 1), self.out_channels, use_bias=True,
                                       activation=self.activation, ndim=self.ndim, device=self.device)
            cur = self.out_channels
        self.conv_out = nn.Conv3d(cur, self.out_channels, kernel_size=1, bias=False, device=self.device)
        self.apply(init_weights_for_snn)

    # -- the network as one engine call ---------------------------------------------------------------------
    def engine(self):
        from ..engine import TransSegEngine
        if self._engine is None:
            self._engine = TransSegEngine(self)
        return self._engine

    def __deepcopy__(self, memo):
        import copy
        eng, self._engine = self._engine, None
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        return new

    def forward(self, x):
        image_size = tuple(x.shape[2:])
        if x.is_meta:
            y = self.conv1(self.conv_in(x) if self.use_resize else x)
            tensors = [y]
            for layer in self.layers:
                y = layer(y)
                tensors.append(y)
            if self.conv_ds is not None:
                y = self.conv_ds(torch.cat(tensors, 1))
            if self.use_resize:
                y = nn.functional.interpolate(y, size=image_size, mode='trilinear')
            y = self.conv_out(y)
            return torch.softmax(y, 1) if self.output_activation_name == 'softmax' else y
        from ..engine import TransSegEngine
        if TransSegEngine.supports(self):
            return self.engine().forward(x)
        return self.forward_modular(x)

    def forward_modular(self, x):
        """The network composed from the stand-alone modules (dense tensors, one autograd node per op)."""
        image_size = tuple(x.shape[2:])
        if self.conv_ds is not None:
            raise NotImplementedError('hno_b200: deep supervision needs blocks the fused engine supports (Hartley operator with '
                                      'shared weights or Hartley multi-head attention, concat skip)')
        if not self.use_resize:
            raise NotImplementedError('hno_b200: use_resize=False needs blocks the fused engine supports')
        x = self.conv1(self.conv_in(x))
        for layer in self.layers:
            x = layer(x)
        # conv_out has no bias and trilinear weights sum to one: the 1x1x1 conv commutes with the interpolation, so
        # it runs at low resolution and only `out_channels` channels are up-sampled (head_kernels.cu)
        w = self.conv_out.weight
        ll = ops.PointwiseConv.apply(x, None, w.view(w.shape[0], -1), None, 0, False)
        tables = get_interp_tables(tuple(x.shape[2:]), image_size, x.device)
        y = ops.HeadUpsample.apply(ll, tables, 1 if self.output_activation_name == 'softmax' else 0)
        return spatial_padcrop(y, image_size)


class NeuralOperatorSeg(_TransSeg):
    """FNO / FNOSeg / HNOSeg family (reference :356-429 over _TransSeg :255-353)."""

    def __init__(self, in_channels, out_channels, filters, num_transform_blocks, num_modes, transform_type,
                 weights_type='shared', use_resize=True, use_deep_supervision=False, use_bias_conv_branch=False,
                 use_block_skip=True, use_block_concat=True, activation='selu',
                 output_activation: Union[str, callable] = 'softmax', ndim=5, device=None):
        super().__init__()
        assert transform_type in ('Fourier', 'Hartley')
        self._init_common(in_channels, out_channels, filters, num_transform_blocks, use_resize, use_deep_supervision,
                          activation, output_activation, ndim, device)
        self.num_modes = (int(num_modes),) * 3 if np.isscalar(num_modes) else tuple(int(m) for m in num_modes)
        self.transform_type = transform_type
        self.weights_type = weights_type
        self.use_bias_conv_branch = use_bias_conv_branch
        self.use_block_skip = use_block_skip
        self.use_block_concat = use_block_concat
        self.block = partial(NeuralOperatorBlock, num_modes=self.num_modes, transform_type=transform_type,
                             weights_type=weights_type, ndim=ndim, activation=activation, device=device,
                             use_bias_conv_branch=use_bias_conv_branch, use_block_skip=use_block_skip,
                             use_block_concat=use_block_concat)
        self.create_layers()


class HartleyMHASeg(_TransSeg):
    """HartleyMHA architecture (reference :432-508): blocks of Hartley-domain multi-head self-attention; deep supervision on
    by default."""

    def __init__(self, in_channels, out_channels, filters, num_transform_blocks, num_heads, num_modes, patch_size,
                 attention_activation='selu', use_resize=True, use_deep_supervision=True, use_bias_conv_branch=False,
                 use_block_skip=True, use_block_concat=True, activation='selu',
                 output_activation: Union[str, callable] = 'softmax', ndim=5, device=None):
        super().__init__()
        self._init_common(in_channels, out_channels, filters, num_transform_blocks, use_resize, use_deep_supervision,
                          activation, output_activation, ndim, device)
        self.num_heads = num_heads
        self.num_modes = num_modes
        self.patch_size = patch_size
        self.attention_activation = attention_activation
        self.use_bias_conv_branch = use_bias_conv_branch
        self.use_block_skip = use_block_skip
        self.use_block_concat = use_block_concat
        self.block = partial(HartleyMHABlock, num_heads=num_heads, num_modes=num_modes, patch_size=patch_size,
                             attention_activation=attention_activation, ndim=ndim, activation=activation, device=device,
                             use_bias_conv_branch=use_bias_conv_branch, use_block_skip=use_block_skip,
                             use_block_concat=use_block_concat)
        self.create_layers()
