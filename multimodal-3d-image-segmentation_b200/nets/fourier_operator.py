"""Fourier spectral layer on the truncated-Hartley kernels (reference: nets/fourier_operator.py:15-223).

The reference computes ``rfftn(norm='forward')`` -> keeps the corners ``[:m0|-m0:] x [:m1|-m1:] x [:m2]`` of the
half-spectrum -> complex channel mixing -> zero-pad -> ``irfftn(s=(-1,-1,W), norm='forward')``.  For real input the
Fourier and Hartley coefficients carry the same information,

    Re F(k) = (H(k) + H(-k)) / 2,        Im F(k) = (H(-k) - H(k)) / 2,          H = (1/N) DHT(x),

and the c2r inverse of a zero-padded half-spectrum ``A + iB`` is ``sum_k c_k (A_k cos(th_k) - B_k sin(th_k))`` with
``c_k = 1`` on the ``k_w = 0`` plane and 2 elsewhere, i.e. an unnormalised inverse DHT of

    H'(k) += c_k (A_k - B_k) / 2,        H'(-k) += c_k (A_k + B_k) / 2.

So the layer is ONE truncated DHT onto the symmetric set ``S = K u (-K)`` (the same tensor-core contraction kernels as
HNOSeg-XS), index algebra + the 24x24 complex mix on the few-MB mode tensor (as two real pointwise convolutions over
the virtual concat [Re; Im]), and ONE adjoint truncated DHT.  The full spectrum and complex tensors are never formed.
"""
import math

import numpy as np
import torch
from torch.nn import Module, Parameter, init

from .. import ops
from ..plan import get_dht_plan


def _axis_sets(n, m, half):
    """Retained frequencies K (reference order), the symmetric set S and the positions of k / -k inside S."""
    K = list(range(m)) if half else list(range(m)) + list(range(n - m, n))
    K = list(dict.fromkeys(K))  # n == 2m: the two corners meet, keep every frequency once
    S = sorted(set(K) | {(-k) % n for k in K})
    pos = {k: i for i, k in enumerate(S)}
    return K, S, [pos[k] for k in K], [pos[(-k) % n] for k in K]


class FourierOperator(Module):
    """Complex channel mixing of the retained Fourier modes; 3-D, with transform; 'shared' (O, I) weights or
    'individual' per-mode weights (O, I, 2 m0, 2 m1, m2) as in experiments/config_files/config_fno.ini."""

    def __init__(self, in_channels, out_channels, num_modes=None, use_bias=False, weights_type='shared',
                 use_transform=True, ndim=5, device=None, dtype=None):
        super().__init__()
        valid = {'individual', 'shared'}
        if weights_type not in valid:
            raise ValueError(f'weights_type must be one of {valid}')
        if ndim != 5:
            raise NotImplementedError('hno_b200 FourierOperator supports 3-D (ndim=5) only')
        if not use_transform:
            raise NotImplementedError('hno_b200 FourierOperator(use_transform=False) (complex inputs) is not supported')
        if use_bias:
            raise NotImplementedError('hno_b200 FourierOperator(use_bias=True) adds to the zero-padded spectrum and is '
                                      'not supported (no reference architecture enables it)')
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
        if weights_type == 'individual':  # rfft keeps the non-negative frequencies of the last axis only (:67-72)
            assert self.num_modes is not None
            shape = shape + tuple(2 * int(m) for m in self.num_modes[:-1]) + (int(self.num_modes[-1]),)
        self.weight_real = Parameter(torch.empty(shape, device=device, dtype=dtype))
        self.weight_imag = Parameter(torch.empty(shape, device=device, dtype=dtype))
        self.register_parameter('bias', None)
        self._geom_cache = {}
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight_real, a=math.sqrt(5))
        init.kaiming_uniform_(self.weight_imag, a=math.sqrt(5))

    def _geometry(self, spatial, device):
        key = (tuple(spatial), str(device))
        g = self._geom_cache.get(key)
        if g is None:
            if self.weights_type == 'individual':  # reference :159: no clamping
                modes = [int(m) for m in self.num_modes]
                assert all(s >= 2 * m for m, s in zip(modes, spatial)), \
                    f'individual weights need a grid of at least twice the modes {modes}, got {tuple(spatial)}'
            else:
                modes = [s // 2 if 2 * int(m) > s else int(m) for m, s in zip(self.num_modes, spatial)]
            axes = [_axis_sets(n, m, half=(a == 2)) for a, (n, m) in enumerate(zip(spatial, modes))]
            plan = get_dht_plan(tuple(spatial), [ax[1] for ax in axes], device)
            ls = [len(ax[1]) for ax in axes]

            def lin(sel):  # linear index into the flattened S grid of the K grid entries
                d = torch.tensor(axes[0][sel]).view(-1, 1, 1)
                h = torch.tensor(axes[1][sel]).view(1, -1, 1)
                w = torch.tensor(axes[2][sel]).view(1, 1, -1)
                return ((d * ls[1] + h) * ls[2] + w).reshape(-1).to(device)

            kshape = tuple(len(ax[0]) for ax in axes)
            ck = torch.full((kshape[2],), 2.0)
            ck[[i for i, k in enumerate(axes[2][0]) if k == 0]] = 1.0  # the k_w = 0 plane counts once in the c2r inverse
            lk, ln = lin(2), lin(3)
            ck5 = ck.view(1, 1, 1, 1, -1)
            # flat tables of the fused mode-domain kernel (hno_fourier_mix_*): int32 positions of k / N - k in S, c_k per K entry
            fused = (lk.to(torch.int32), ln.to(torch.int32),
                     ck5.expand((1, 1) + kshape).reshape(-1).contiguous().to(device=device, dtype=torch.float32))
            g = self._geom_cache[key] = (plan, lk, ln, kshape, tuple(ls), ck5.to(device), fused)
        return g

    def spectral(self, x):
        """x -> (H', plan): the Hartley coefficients on S whose unnormalised inverse DHT is the layer's output."""
        x = x.contiguous()
        B = x.shape[0]
        plan, lin_k, lin_n, kshape, ls, ck, fused = self._geometry(tuple(x.shape[2:]), x.device)
        z = ops.TruncatedDHT.apply(x, plan).reshape(B, self.in_channels, -1)  # (1/N) DHT on S
        if self.in_channels % 4 == 0 and self.out_channels % 4 == 0:
            # one kernel per direction instead of index_select x 2 + elementwise + two channel mixes + index_add x 2
            hp = ops.FourierMixShared.apply(z, self.weight_real, self.weight_imag, *fused)
            return hp.reshape((B, self.out_channels) + ls), plan
        hk = z.index_select(2, lin_k).reshape((B, self.in_channels) + kshape)
        hn = z.index_select(2, lin_n).reshape((B, self.in_channels) + kshape)
        re = ((hk + hn) * 0.5).contiguous()
        im = ((hn - hk) * 0.5).contiguous()
        wr, wi = self.weight_real, self.weight_imag
        if self.weights_type == 'individual':
            a, b = ops.ComplexModeMix.apply(re, im, wr, wi)
        else:
            a = ops.PointwiseConv.apply(re, im, torch.cat([wr, -wi], 1), None, 0, False)  # Re of (wr + i wi)(re + i im)
            b = ops.PointwiseConv.apply(re, im, torch.cat([wi, wr], 1), None, 0, False)   # Im
        hp = x.new_zeros((B, self.out_channels, ls[0] * ls[1] * ls[2]))
        hp = hp.index_add(2, lin_k, (ck * (a - b) * 0.5).reshape(B, self.out_channels, -1))
        hp = hp.index_add(2, lin_n, (ck * (a + b) * 0.5).reshape(B, self.out_channels, -1))
        return hp.reshape((B, self.out_channels) + ls), plan

    def forward(self, inputs):
        if inputs.is_meta:
            return inputs.new_empty((inputs.shape[0], self.out_channels) + tuple(inputs.shape[2:]))
        hp, plan = self.spectral(inputs)
        return ops.TruncatedIDHT.apply(hp, plan)
