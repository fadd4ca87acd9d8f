"""HNOSeg-XS on the B200 kernels (reference: nets/hnosegxs.py).

Module tree, constructor signatures and ``state_dict`` keys follow the reference so checkpoints are
interchangeable; ``forward`` does not: the whole network runs through ``engine.XSEngine`` (one autograd node,
planar padded activations, fused epilogues).  The sub-modules remain usable on their own with ordinary dense
NCDHW tensors.
"""
from functools import partial
from typing import Union

import numpy as np
import torch
from torch import nn

from .. import ops
from ..plan import get_crop_plan, get_dht_plan
from .hartley_operator import HartleyOperator
from .nets_utils import ConvNormAct, _is_selu, init_weights_for_snn, spatial_padcrop


def _modes_tuple(num_modes, ndim):
    if np.isscalar(num_modes):
        return (int(num_modes),) * (ndim - 2)
    assert len(num_modes) == ndim - 2
    return tuple(int(m) for m in num_modes)


class TransformCrop(nn.Module):
    """DHT of the trailing axes restricted to the low/high corners (reference :332-410), as ONE truncated
    contraction: the full spectrum is never formed."""

    def __init__(self, num_modes, ndim):
        super().__init__()
        assert ndim in (4, 5)
        self.num_modes = _modes_tuple(num_modes, ndim)

    def forward(self, x):
        spatial = tuple(x.shape[2:])
        assert len(spatial) == len(self.num_modes)
        modes = tuple(s // 2 if 2 * m > s else m for m, s in zip(self.num_modes, spatial))
        if x.is_meta:
            return x.new_empty(tuple(x.shape[:2]) + tuple(2 * m for m in modes))
        if x.ndim == 4:  # 2-D: a 3-D problem with a singleton leading axis
            z = ops.TruncatedDHT.apply(x.unsqueeze(2), get_dht_plan((1,) + spatial, [[0]] + [
                list(range(m)) + list(range(n - m, n)) for n, m in zip(spatial, modes)], x.device))
            return z.squeeze(2)
        return ops.TruncatedDHT.apply(x, get_crop_plan(spatial, modes, x.device))


class PadInverse(nn.Module):
    """Zero-pad the corner modes to `spatial_shape` and take the unnormalised inverse DHT (reference :413-494),
    evaluated as the adjoint truncated contraction: the padded spectrum is never formed."""

    def __init__(self, ndim):
        super().__init__()
        assert ndim in (4, 5)

    def forward(self, x, spatial_shape):
        spatial = tuple(int(s) for s in spatial_shape)
        modes = tuple(s // 2 for s in x.shape[2:])
        assert all(n >= 2 * m for n, m in zip(spatial, modes))
        if x.is_meta:
            return x.new_empty(tuple(x.shape[:2]) + spatial)
        if tuple(x.shape[2:]) != tuple(2 * m for m in modes):
            raise ValueError('PadInverse expects an even number of retained modes per axis')
        kl = [list(range(m)) + list(range(n - m, n)) for n, m in zip(spatial, modes)]
        if x.ndim == 4:
            y = ops.TruncatedIDHT.apply(x.unsqueeze(2), get_dht_plan((1,) + spatial, [[0]] + kl, x.device))
            return y.squeeze(2)
        return ops.TruncatedIDHT.apply(x, get_dht_plan(spatial, kl, x.device))


class NeuralOperatorBlock(nn.Module):
    """One frequency-domain convolution: selu(op(x) + x) in a single kernel (reference :282-329)."""

    def __init__(self, in_channels, out_channels, num_modes, weights_type, ndim, activation, device,
                 use_conv_branch=False):
        super().__init__()
        if use_conv_branch:
            raise NotImplementedError('hno_b200: use_conv_branch is not used by HNOSeg-XS and is not supported')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 implements the SELU (self-normalising) variant only')
        if in_channels != out_channels:
            raise NotImplementedError('hno_b200: the residual mode mix needs in_channels == out_channels')
        self.op = HartleyOperator(in_channels, out_channels, num_modes, use_bias=False, weights_type=weights_type,
                                  use_transform=False, ndim=ndim, device=device)
        self.conv_branch = None
        self.normalization = None
        self.activation = nn.functional.selu

    def forward(self, x):
        if x.is_meta:
            return torch.empty_like(x)
        return self.op._mix(x, act=1, residual=True)


class HNOXSBlock(nn.Module):
    """HNO-XS block (reference :185-279): [mapping conv] -> DHT+crop -> n_XS mixes -> pad+inverse DHT -> SELU ->
    concat skip + 1x1 conv + SELU."""

    def __init__(self, num_convs, in_channels, out_channels, num_modes, weights_type='shared', ndim=5,
                 activation='selu', device=None, use_conv_branch=False, use_block_concat=True):
        super().__init__()
        if ndim != 5:
            raise NotImplementedError('hno_b200 HNOXSBlock supports 3-D (ndim=5) only')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 implements the SELU (self-normalising) variant only')
        cur = in_channels
        self.mapping_conv = None
        if cur != out_channels:
            self.mapping_conv = ConvNormAct(cur, out_channels, use_bias=True, activation=activation, ndim=ndim,
                                            device=device)
            cur = out_channels
        self.transform_crop = TransformCrop(num_modes, ndim)
        self.conv_blocks = nn.ModuleList()
        for _ in range(num_convs):
            self.conv_blocks.append(NeuralOperatorBlock(cur, out_channels, num_modes, weights_type, ndim, activation,
                                                        device, use_conv_branch))
            cur = out_channels
        self.pad_inverse = PadInverse(ndim)
        self.normalization = None
        self.activation = nn.functional.selu
        self.conv_concat = None
        if use_block_concat:
            self.conv_concat = ConvNormAct(cur + out_channels, out_channels, use_bias=True, activation=activation,
                                           ndim=ndim, device=device)

    def forward(self, x):
        if self.mapping_conv is not None:
            x = self.mapping_conv(x)
        if x.is_meta:
            return torch.empty_like(x)
        skip = x
        spatial = tuple(x.shape[2:])
        z = self.transform_crop(x)
        for block in self.conv_blocks:
            z = block(z)
        y = self.activation(self.pad_inverse(z, spatial))
        if self.conv_concat is not None:
            op = self.conv_concat.op
            return ops.PointwiseConv.apply(y, skip, op.weight, op.bias, 1, False)  # virtual torch.cat([y, skip], 1)
        return y + skip


class HNOSegXS(nn.Module):
    """HNOSeg-XS (reference :20-182).  Same arguments; 3-D, SELU, softmax/identity output on the CUDA path."""

    def __init__(self, in_channels, out_channels, filters, num_transform_blocks, num_modes, weights_type='shared',
                 use_resize=True, use_deep_supervision=False, use_unet_skip=True, use_block_concat=True,
                 activation='selu', output_activation: Union[str, callable] = 'softmax', ndim=5, device=None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.filters = filters
        self.num_transform_blocks = num_transform_blocks
        self.num_modes = num_modes
        self.weights_type = weights_type
        self.use_resize = use_resize
        self.use_deep_supervision = use_deep_supervision
        self.use_unet_skip = use_unet_skip
        self.use_block_concat = use_block_concat
        self.activation = activation
        self.output_activation = output_activation
        self.ndim = ndim
        self.device = device
        assert self.ndim in (4, 5)
        if self.ndim != 5:
            raise NotImplementedError('hno_b200 HNOSegXS supports 3-D volumes (ndim=5) only')
        if not _is_selu(activation):
            raise NotImplementedError('hno_b200 HNOSegXS implements the SELU (self-normalising) variant only')
        if output_activation not in ('softmax', None, 'identity'):
            raise NotImplementedError("hno_b200 HNOSegXS supports output_activation in {'softmax', None}")
        if np.isscalar(self.num_transform_blocks):
            self.num_transform_blocks = [self.num_transform_blocks]
        self.num_modes = _modes_tuple(num_modes, ndim)
        self._engine = None
        self.create_layers()

    def create_layers(self):
        block = partial(HNOXSBlock, num_modes=self.num_modes, weights_type=self.weights_type, ndim=self.ndim,
                        activation=self.activation, device=self.device, use_block_concat=self.use_block_concat)
        f = self.filters
        self.conv_in = None
        cur = self.in_channels
        if self.use_resize:  # reference :102-105; without it the blocks run at the image resolution
            self.conv_in = ConvNormAct(cur, f, kernel_size=2, stride=2, use_bias=True, activation=self.activation,
                                       ndim=self.ndim, device=self.device)
            cur = f
        self.conv1 = ConvNormAct(cur, f, use_bias=True, activation=self.activation, ndim=self.ndim, device=self.device)
        self.layers = nn.ModuleList()
        nb = len(self.num_transform_blocks)
        for i, n_convs in enumerate(self.num_transform_blocks):
            cin = f + (f if (self.use_unet_skip and i > nb // 2) else 0)
            self.layers.append(block(n_convs, cin, f))
        # deep supervision (reference :110-125, 134): conv_out reads the concatenation of conv1's and every block's output
        cout_in = f * (nb + 1) if self.use_deep_supervision else f
        self.conv_out = nn.Conv3d(cout_in, self.out_channels, kernel_size=1, bias=False, device=self.device)
        self.apply(init_weights_for_snn)

    # -- the network as one engine call ---------------------------------------------------------------------
    def engine(self):
        from ..engine import XSEngine
        if self._engine is None:
            self._engine = XSEngine(self)
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
        if x.is_meta:  # torchinfo / torchview trace of experiments/train_test.py:118-122
            return x.new_empty((x.shape[0], self.out_channels) + tuple(x.shape[2:]))
        return self.engine().forward(x)

    def forward_logits(self, x):
        """Full-resolution logits, i.e. the output of conv_out (reference :178) before the softmax."""
        with torch.no_grad():
            _, S = self.engine().run_forward(x, save=False, head=False)
            if S.tables is None:  # use_resize=False
                return ops.head_direct_forward(S.ll, S.geom[:3], 0)
            return ops.head_forward(S.ll, S.tables, S.geom[3], 0)

    def predict_labels(self, x):
        """Inference as experiments/train_test.py:398-408 uses the model (forward under no_grad, argmax over the classes),
        with the argmax on the device: returns the uint8 label map (B, D, H, W).  The probabilities are never
        materialised (4 bytes x classes per voxel neither written nor copied back)."""
        from ..engine import preferred_axis_perm
        with torch.no_grad():
            # a large volume whose last axis is not its shortest (SimpleITK's (z, y, x) order) runs on permuted axes and the
            # one-byte label map is transposed back (engine.preferred_axis_perm)
            perm = preferred_axis_perm(self, tuple(x.shape[2:])) if x.ndim == 5 else None
            _, S = self.engine().run_forward(x, save=False, head=False, perm=perm)
            if S.tables is None:  # use_resize=False
                return ops.head_direct_argmax(S.ll, S.geom[:3])
            labels = ops.head_argmax(S.ll, S.tables, S.geom[3])
            if perm is not None:
                labels = ops.permute_spatial(labels, tuple(perm.index(i) for i in range(3)))
            return labels

    def loss(self, x, labels, loss_name='DiceLoss', param=None):
        """Fused head + loss on integer labels; equals loss_fn(self(x), to_categorical(labels)) for loss_name in
        DiceLoss / PCCLoss / ExpDiceLoss(exp=param) / CrossEntropyLoss."""
        return self.engine().loss(x, labels, loss_name, param)

    def forward_modular(self, x):
        """Same network composed from the stand-alone sub-modules (dense tensors, one autograd node per op).
        Slower than forward(); kept as an independent cross-check of the engine."""
        if self.use_deep_supervision:
            raise NotImplementedError('hno_b200: forward_modular does not implement deep supervision; use forward()')
        image_size = tuple(x.shape[2:])
        if not self.use_resize:
            raise NotImplementedError('hno_b200: forward_modular covers the use_resize=True network; use forward()')
        x = self.conv1(self.conv_in(x))
        nb = len(self.layers)
        stash = {}
        for i, layer in enumerate(self.layers):
            if layer.mapping_conv is not None:
                op = layer.mapping_conv.op
                x = ops.PointwiseConv.apply(x, stash[nb - 1 - i], op.weight, op.bias, 1, False)
                mc, layer.mapping_conv = layer.mapping_conv, None
                try:
                    x = layer(x)
                finally:
                    layer.mapping_conv = mc
            else:
                x = layer(x)
            if self.use_unet_skip and i < nb // 2:
                stash[i] = x
        low = ops.PointwiseConv.apply(x, None, self.conv_out.weight, None, 0, False)
        tables = ops.get_interp_tables(tuple(low.shape[2:]), image_size, x.device)
        act = 1 if self.output_activation == 'softmax' else 0
        return spatial_padcrop(ops.HeadUpsample.apply(low, tables, act), image_size)
