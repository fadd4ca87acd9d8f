"""Hartley multi-head attention (reference: nets/hartley_mha.py).

Same constructor, parameter names / shapes (``weight_query`` (H, key_dim, C), ``weight_key``, ``weight_value``,
``weight_out`` (value_dim, value_dim * H), optional ``bias_*``) and call conventions (one tensor, or a list of two / three
for cross-attention) as the reference.  The computation is the reference's

    DHT -> per-head 1x1x1 Q/K/V on the 8 corners -> patch grouping -> att = selu(Q^T K / sqrt(c)) -> att V -> ungrouping ->
    output projection -> zero-pad + inverse DHT                                          (hartley_mha.py:136-222)

on the CUDA kernels: ONE truncated transform per distinct input (hno_dht3_forward; the full spectrum is never formed),
the projections / grouping written straight into the attention layouts (csrc/mha_kernels.cu), the two big contractions
as tcgen05 3xTF32 GEMMs with the SELU in the epilogue (csrc/gemm_tc.cu), and the adjoint truncated transform.
"""
import math
from typing import Union

import numpy as np
import torch
from torch.nn import Module, Parameter, init

from .. import ops
from ..plan import get_dht_plan


def _activation_code(act):
    if act is None:
        return 0
    if act == 'selu' or act is torch.nn.functional.selu:
        return 1
    raise NotImplementedError("hno_b200 HartleyMultiHeadAttention supports attention_activation 'selu' or None")


class HartleyMultiHeadAttention(Module):
    def __init__(self, in_channels, key_dim, num_heads, num_modes, patch_size=None,
                 attention_activation: Union[str, callable] = 'selu', value_dim=None, key_in_channels=None,
                 value_in_channels=None, use_bias=False, use_transform=True, ndim=5, device=None, dtype=None):
        super().__init__()
        if ndim != 5:
            raise NotImplementedError('hno_b200 HartleyMultiHeadAttention supports 3-D (ndim=5) only')
        kw = {'device': device, 'dtype': dtype}
        self.in_channels = in_channels
        self.key_dim = key_dim
        self.num_heads = num_heads
        self.num_modes = num_modes
        self.patch_size = patch_size
        self.attention_activation = attention_activation
        self._act = _activation_code(attention_activation)
        self.value_dim = value_dim or key_dim
        self.key_in_channels = key_in_channels or in_channels
        self.value_in_channels = value_in_channels or self.key_in_channels
        self.use_bias = use_bias
        self.use_transform = use_transform
        if np.isscalar(self.num_modes):
            self.num_modes = (self.num_modes,) * (ndim - 2)
        else:
            assert len(self.num_modes) == ndim - 2
            self.num_modes = tuple(self.num_modes)
        if np.isscalar(self.patch_size):
            self.patch_size = (self.patch_size,) * (ndim - 2)
        if isinstance(self.attention_activation, str):
            self.attention_activation = getattr(torch.nn.functional, self.attention_activation)
        self.weight_query = Parameter(torch.empty((num_heads, key_dim, in_channels), **kw))
        self.weight_key = Parameter(torch.empty((num_heads, key_dim, self.key_in_channels), **kw))
        self.weight_value = Parameter(torch.empty((num_heads, self.value_dim, self.value_in_channels), **kw))
        self.weight_out = Parameter(torch.empty((self.value_dim, self.value_dim * num_heads), **kw))
        if use_bias:
            ones = (1,) * (ndim - 2)
            self.bias_query = Parameter(torch.empty((1, num_heads, key_dim) + ones, **kw))
            self.bias_key = Parameter(torch.empty((1, num_heads, key_dim) + ones, **kw))
            self.bias_value = Parameter(torch.empty((1, num_heads, self.value_dim) + ones, **kw))
            self.bias_out = Parameter(torch.empty((1, self.value_dim) + ones, **kw))
        else:
            for name in ('bias_query', 'bias_key', 'bias_value', 'bias_out'):
                self.register_parameter(name, None)
        self.reset_parameters()

    def reset_parameters(self):
        for w, b in ((self.weight_query, self.bias_query), (self.weight_key, self.bias_key),
                     (self.weight_value, self.bias_value), (self.weight_out, self.bias_out)):
            init.kaiming_uniform_(w, a=math.sqrt(5))
            if b is not None:
                init.zeros_(b)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _split_inputs(inputs):
        if not isinstance(inputs, (tuple, list)):
            return inputs, None, None
        if len(inputs) == 2:
            return inputs[0], inputs[1], None
        if len(inputs) == 3:
            return inputs[0], inputs[1], inputs[2]
        raise ValueError('Invalid inputs.')

    def _plan(self, spatial, device):
        assert all(s >= 2 * m for s, m in zip(spatial, self.num_modes))  # reference :165-172
        kl = [list(range(m)) + list(range(n - m, n)) for n, m in zip(spatial, self.num_modes)]
        return get_dht_plan(spatial, kl, device)

    def attend(self, zq, zk=None, zv=None):
        """Retained modes in -> retained modes out (the reference's _call_notransform, :224-296)."""
        def flat(b):
            return None if b is None else b.reshape(b.shape[1], b.shape[2]).contiguous()
        bo = None if self.bias_out is None else self.bias_out.reshape(-1).contiguous()
        return ops.HartleyAttention.apply(zq, zk, zv, self.weight_query, self.weight_key, self.weight_value, self.weight_out,
                                          flat(self.bias_query), flat(self.bias_key), flat(self.bias_value), bo,
                                          self.patch_size, self._act)

    def spectral(self, inputs):
        """inputs -> (modes of the layer output, plan whose adjoint transform gives the output)."""
        q, k, v = self._split_inputs(inputs)
        spatial = tuple(q.shape[2:])
        plan = self._plan(spatial, q.device)
        dht = lambda t: ops.TruncatedDHT.apply(t.contiguous(), plan)  # noqa: E731
        return self.attend(dht(q), None if k is None else dht(k), None if v is None else dht(v)), plan

    def forward(self, inputs):
        q, k, v = self._split_inputs(inputs)
        if q.is_meta:
            shape = tuple(q.shape[2:])
            return q.new_empty((q.shape[0], self.value_dim) + shape)
        if not self.use_transform:
            return self.attend(q.contiguous(), None if k is None else k.contiguous(), None if v is None else v.contiguous())
        z, plan = self.spectral(inputs)
        return ops.TruncatedIDHT.apply(z, plan)
