"""Hartley spectral layer (reference: nets/hartley_operator.py)."""
import math

import numpy as np
import torch
from torch.nn import Module, Parameter, init

from .. import ops
from ..plan import get_crop_plan, get_dht_plan


class HartleyOperator(Module):
    """Channel mixing of retained Hartley modes.

    ``use_transform=False`` (HNOSeg-XS): the input already holds the cropped modes; 'shared' weights are one
    (O, I) matrix for all modes, 'individual' weights are (O, I, 2m0, 2m1, 2m2) and use the Hartley even/odd
    recombination.  ``use_transform=True`` (HNOSeg): truncated DHT -> mix -> SELU -> adjoint DHT, which equals
    the reference's transform / 8-corner mix / zero-pad / SELU / inverse chain because selu(0) == 0.
    """

    def __init__(self, in_channels, out_channels, num_modes=None, use_bias=False, weights_type='shared',
                 use_transform=True, ndim=5, device=None, dtype=None):
        super().__init__()
        valid = {'individual', 'shared'}
        if weights_type not in valid:
            raise ValueError(f'weights_type must be one of {valid}')
        if ndim != 5:
            raise NotImplementedError('hno_b200 HartleyOperator supports 3-D (ndim=5) only')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_modes = num_modes
        self.use_bias = use_bias
        self.weights_type = weights_type
        self.use_transform = use_transform
        if self.num_modes is not None:
            if np.isscalar(self.num_modes):
                self.num_modes = (self.num_modes,) * (ndim - 2)
            else:
                assert len(self.num_modes) == ndim - 2
                self.num_modes = tuple(self.num_modes)
        shape = (out_channels, in_channels)
        if weights_type == 'individual':
            assert self.num_modes is not None
            shape = shape + tuple(2 * int(m) for m in self.num_modes)
        self.weight = Parameter(torch.empty(shape, device=device, dtype=dtype))
        if use_bias:
            self.bias = Parameter(torch.empty((1, out_channels) + (1,) * (ndim - 2), device=device, dtype=dtype))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            init.zeros_(self.bias)

    def _mix(self, z, act=0, residual=False):
        if self.weights_type == 'shared':
            return ops.PointwiseConv.apply(z, None, self.weight, None, act, residual)
        assert not (act and not residual)
        return ops.HartleyConv.apply(z, self.weight, residual)

    def forward(self, inputs):
        if inputs.is_meta:
            return inputs.new_empty((inputs.shape[0], self.out_channels) + tuple(inputs.shape[2:]))
        if not self.use_transform:
            x = self._mix(inputs)
            if self.use_bias:
                x = x + self.bias
            return x
        if self.use_bias:
            raise NotImplementedError('hno_b200: HartleyOperator(use_transform=True, use_bias=True) is not supported '
                                      '(no reference architecture enables it)')
        z, plan = self.spectral(inputs)
        return ops.TruncatedIDHT.apply(z, plan)

    def spectral(self, inputs):
        """inputs -> (z, plan): the mixed, SELU-activated retained modes and the plan whose adjoint transform gives the
        layer output (the HNOSeg block fuses that adjoint with its conv branch, nets/architectures.py)."""
        spatial = tuple(inputs.shape[2:])
        if self.weights_type == 'shared':
            plan = get_crop_plan(spatial, self.num_modes, inputs.device)
            z = ops.TruncatedDHT.apply(inputs, plan)
            return ops.PointwiseConv.apply(z, None, self.weight, None, 1, False), plan  # mix + SELU on the modes
        # individual weights on the FULL spectrum's reversal (reference :196-241): the partner of the retained
        # index n-m is +m, which lies outside the retained set, so the forward transform also produces it.
        m = [int(v) for v in self.num_modes]
        assert all(n >= 2 * mm for n, mm in zip(spatial, m))
        kl = [list(range(mm)) + list(range(n - mm, n)) for n, mm in zip(spatial, m)]
        ext = [k + ([mm] if mm not in k else []) for k, mm in zip(kl, m)]
        z_ext = ops.TruncatedDHT.apply(inputs, get_dht_plan(spatial, ext, inputs.device))
        z = ops.HartleyConvFull.apply(z_ext, self.weight, _partner_table(spatial, tuple(m), inputs.device), 1)
        return z, get_dht_plan(spatial, kl, inputs.device)


_partner_tables = {}


def _partner_table(spatial, m, device):
    """int32 device tensor [r0 | r1 | r2]: position, in the extended mode list (retained + [+m]), of the full-spectrum
    reversal partner (n - k) mod n of every retained frequency k (reference get_reverse on the N-point spectrum, :199)."""
    key = (tuple(spatial), tuple(m), str(device))
    t = _partner_tables.get(key)
    if t is None:
        tab = []
        for n, mm in zip(spatial, m):
            ks = list(range(mm)) + list(range(n - mm, n))
            pos = {k: i for i, k in enumerate(ks)}
            pos.setdefault(mm, len(ks))
            tab += [pos[(n - k) % n] for k in ks]
        t = _partner_tables[key] = torch.tensor(tab, dtype=torch.int32, device=device)
    return t


def hartley_conv(equation, weight, weight_reverse, x, x_reverse):
    """Functional form kept for API compatibility (reference :302-317); plain torch, used off the hot path."""
    h1 = torch.einsum(equation, weight, x + x_reverse)
    h2 = torch.einsum(equation, weight_reverse, x - x_reverse)
    return (h1 + h2) * 0.5


def get_reverse(x, dims):
    """x[(N - k) mod N] along `dims` (reference :320-333)."""
    assert isinstance(dims, (list, tuple))
    return torch.roll(torch.flip(x, dims), [1] * len(dims), dims)
