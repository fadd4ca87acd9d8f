"""Discrete Hartley transforms on the GPU without an FFT (reference: nets/dht.py:16-66).

``dhtn(x, dim, is_inverse)`` keeps the reference signature.  The full spectrum is the special case
"retain every frequency" of the truncated cas-basis contraction in csrc/dht_kernels.cu; HNOSeg-XS never
needs it (TransformCrop / PadInverse contract straight to / from the retained corners), it exists for API
completeness and costs O(N * (D + H + W)) per channel instead of an FFT's O(N log N).
"""
import torch

from .. import ops
from ..plan import get_dht_plan


def _as_5d(x, dim):
    nd = x.ndim
    dims = sorted(d % nd for d in dim)
    k = len(dims)
    if k not in (2, 3) or dims != list(range(nd - k, nd)):
        raise NotImplementedError('hno_b200 dhtn transforms the trailing 2 or 3 dimensions only')
    spatial = tuple(x.shape[nd - k:])
    lead = x.shape[:nd - k]
    if k == 2:
        spatial = (1,) + spatial
    return x.reshape((-1, 1) + spatial), spatial, lead


def dhtn(x, dim, is_inverse=False):
    """(Inverse) DHT over the trailing dimensions: 1/N-normalised forward, unnormalised inverse."""
    if x.is_meta:
        return torch.empty_like(x)
    x5, spatial, _ = _as_5d(x, dim)
    plan = get_dht_plan(spatial, [list(range(n)) for n in spatial], x.device)
    y = (ops.TruncatedIDHT if is_inverse else ops.TruncatedDHT).apply(x5.contiguous(), plan)
    return y.reshape(x.shape)


def dht2(x, is_inverse=False):
    return dhtn(x, dim=(-2, -1), is_inverse=is_inverse)


def dht3(x, is_inverse=False):
    return dhtn(x, dim=(-3, -2, -1), is_inverse=is_inverse)
