"""Host-side plans: DHT table blobs and interpolation tables, built by the C library and cached per device."""
import ctypes

import numpy as np
import torch

from . import _lib

_INT_P = ctypes.POINTER(ctypes.c_int)


def corner_frequencies(n, m):
    """Retained frequencies of TransformCrop on one axis (reference nets/hnosegxs.py:382-410): the mode count is
    clamped to n // 2 when it does not fit twice, low block first, then the high block."""
    m = int(m)
    if 2 * m > n:
        m = n // 2
    return list(range(m)) + list(range(n - m, n))


def plane_pitch(h, w):
    """Plane pitch (floats) used for the model's internal activations: H*W rounded up to 32 floats (128 B)."""
    return (h * w + 31) // 32 * 32


class DhtPlan:
    """Table blob for the truncated 3-D DHT on a (D, H, W) grid with per-axis frequency lists."""

    def __init__(self, spatial, klists, device=None):
        lib = _lib.load()
        D, H, W = (int(s) for s in spatial)
        ks = [np.ascontiguousarray(np.asarray(k, dtype=np.int32)) for k in klists]
        L = [int(k.size) for k in ks]
        nbytes = lib.hno_dht3_plan_bytes(D, H, W, *L)
        host = torch.zeros((nbytes + 3) // 4, dtype=torch.int32)
        _lib.call('hno_dht3_plan_fill', host.data_ptr(), nbytes, D, H, W,
                  ks[0].ctypes.data_as(ctypes.c_void_p), L[0], ks[1].ctypes.data_as(ctypes.c_void_p), L[1],
                  ks[2].ctypes.data_as(ctypes.c_void_p), L[2])
        words = int(host[2])
        self.host = host[:words].contiguous()
        self.spatial = (D, H, W)
        self.klists = [k.tolist() for k in ks]
        self.modes_shape = tuple(L)
        hdr = self.host.numpy()
        self.axes = []
        for a in range(3):
            f = hdr[4 + 16 * a: 4 + 16 * (a + 1)]
            self.axes.append(dict(n=int(f[0]), L=int(f[1]), JC=int(f[2]), JS=int(f[3]), J=int(f[4]), nh=int(f[5]),
                                  JCp=int(f[6]), JSp=int(f[7]), off_fcos=int(f[8]), off_fsin=int(f[9]),
                                  off_full=int(f[10]), off_kdesc=int(f[11]), off_jdesc=int(f[12])))
        self.dev = self.host.to(device) if device is not None and torch.device(device).type == 'cuda' else None

    @property
    def n_voxels(self):
        return self.spatial[0] * self.spatial[1] * self.spatial[2]

    def workspace_bytes(self, pitch, nslab):
        return int(_lib.load().hno_dht3_workspace_bytes(self.host.data_ptr(), int(pitch), int(nslab)))

    # views used by the CPU emulation in tests (never by the product path)
    def table(self, axis, name):
        ax = self.axes[axis]
        h = self.host.numpy()
        if name == 'full':
            return h[ax['off_full']: ax['off_full'] + ax['J'] * ax['n']].view(np.float32).reshape(ax['J'], ax['n'])
        if name == 'fcos':
            return h[ax['off_fcos']: ax['off_fcos'] + (ax['nh'] + 1) * ax['JCp']].view(np.float32).reshape(-1, ax['JCp'])
        if name == 'fsin':
            return h[ax['off_fsin']: ax['off_fsin'] + (ax['nh'] + 1) * ax['JSp']].view(np.float32).reshape(-1, ax['JSp'])
        if name == 'kdesc':
            return h[ax['off_kdesc']: ax['off_kdesc'] + 4 * ax['L']].reshape(ax['L'], 4)
        if name == 'jdesc':
            return h[ax['off_jdesc']: ax['off_jdesc'] + 4 * ax['J']].reshape(ax['J'], 4)
        raise KeyError(name)


class InterpTables:
    """Per-axis trilinear source indices / weights (low res (D,H,W) -> high res (Dx,Hx,Wx)) and their inverse ranges."""

    def __init__(self, lo, hi, device=None):
        lib = _lib.load()
        lo = tuple(int(v) for v in lo)
        hi = tuple(int(v) for v in hi)
        nbytes = lib.hno_interp_tables_bytes(*lo, *hi)
        self.host = torch.zeros((nbytes + 3) // 4, dtype=torch.int32)
        _lib.call('hno_interp_tables_fill', self.host.data_ptr(), nbytes, *lo, *hi)
        self.lo, self.hi = lo, hi
        self.dev = self.host.to(device) if device is not None and torch.device(device).type == 'cuda' else None

    def axis(self, a):
        """(i0, i1, lambda1, start, end) numpy views of one axis, for tests."""
        h = self.host.numpy()
        off = lambda base: int(h[7 + base * 3 + a])  # noqa: E731  header: magic, lo[3], hi[3], then 5 offset triples
        nh, nl = self.hi[a], self.lo[a]
        return (h[off(0): off(0) + nh], h[off(1): off(1) + nh], h[off(2): off(2) + nh].view(np.float32),
                h[off(3): off(3) + nl], h[off(4): off(4) + nl])

    def head_backward_workspace_bytes(self, B, C):
        return int(_lib.load().hno_head_backward_workspace_bytes(self.host.data_ptr(), int(B), int(C)))


_plan_cache = {}
_interp_cache = {}


def get_dht_plan(spatial, klists, device):
    key = (tuple(int(s) for s in spatial), tuple(tuple(int(v) for v in k) for k in klists), str(device))
    plan = _plan_cache.get(key)
    if plan is None:
        plan = _plan_cache[key] = DhtPlan(spatial, klists, device)
    return plan


def get_crop_plan(spatial, modes, device):
    """Plan of TransformCrop / PadInverse for `modes` (clamped like the reference) on a grid `spatial`."""
    return get_dht_plan(spatial, [corner_frequencies(n, m) for n, m in zip(spatial, modes)], device)


def get_interp_tables(lo, hi, device):
    key = (tuple(int(v) for v in lo), tuple(int(v) for v in hi), str(device))
    t = _interp_cache.get(key)
    if t is None:
        t = _interp_cache[key] = InterpTables(lo, hi, device)
    return t
