"""`ImageTransform` of the reference's experiments/data_io/dataset.py:63-192 for tensors that are already on the GPU.

The reference augments every sample inside the DataLoader workers: numpy generator draws -> homogeneous matrix ->
SimpleITK nearest-neighbour resampling channel by channel (:205-237) -> numpy flips (:240-245).  At hundreds of volumes
per second per GPU that loader is the bottleneck, so here only the PARAMETERS are drawn on the host
-- the same generator, the same draws in the same order, so a seed reproduces the reference's augmentation sequence --
and the image work is one gather kernel over the whole batch (`hno_affine_resample_nn`, csrc/input_kernels.cu) that
resamples and flips images (float32 or raw int16) and their label maps (uint8) alike.

Same constructor arguments as the reference.  No CPU path: a non-CUDA tensor raises.
"""
import numpy as np
import torch

from ..._lib import call, ptr, stream_ptr


def _rotation(angles_xyz):
    """Homogeneous 4 x 4 rotation Rz(c) Ry(b) Rx(a) for angles about SimpleITK's x, y, z axes (dataset.py:133-146)."""
    a, b, c = angles_xyz
    sa, ca, sb, cb, sc, cc = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(c), np.cos(c)
    out = np.eye(4)
    out[0, :3] = (cb * cc, -ca * sc + sa * sb * cc, sa * sc + ca * sb * cc)
    out[1, :3] = (cb * sc, ca * cc + sa * sb * sc, -sa * cc + ca * sb * sc)
    out[2, :3] = (-sb, sa * cb, ca * cb)
    return out


def draw_transform(rng, spatial, rotation_range=None, shift_range=None, zoom_range=None, flip=None,
                   augmentation_probability=1.0):
    """One sample's augmentation parameters, drawn from `rng` exactly as ImageTransform.__call__ does (dataset.py:106-178):
    the binomial gate, one uniform per non-zero rotation entry, one per non-zero shift entry, the zoom, and -- after the
    geometric part -- one random() per enabled flip axis.

    Returns (xform, flags): xform = the 3 x 4 fp64 matrix (SimpleITK (x, y, z) order) from an output index to the
    continuous input index, i.e. what the reference gives sitk.AffineTransform after centring about size / 2 + 0.5
    (:195-202, :220-224), or None when no geometric transform was drawn; flags = bit 0 / 1 / 2 for a flip of the
    first / second / third spatial axis (for 2-D images the H and W bits)."""
    nd = len(spatial)
    if not rng.binomial(1, augmentation_probability):
        return None, 0
    angles = None
    if rotation_range is not None:
        if np.isscalar(rotation_range):
            assert nd == 2
            angles = np.pi / 180 * rng.uniform(-rotation_range, rotation_range) if rotation_range else 0
        else:
            assert len(rotation_range) == 3
            angles = [np.pi / 180 * rng.uniform(-r, r) if r else 0 for r in rotation_range]
    shift = None
    if shift_range is not None:
        assert len(shift_range) == nd
        shift = [rng.uniform(-s, s) * spatial[i] if s else 0 for i, s in enumerate(shift_range)]
    zoom = None
    if zoom_range is not None:
        zoom = rng.uniform(zoom_range[0], zoom_range[1])

    matrix = None
    if angles is not None:
        if np.isscalar(angles):
            if angles != 0:
                matrix = np.eye(3)
                matrix[0, :2] = (np.cos(angles), -np.sin(angles))
                matrix[1, :2] = (np.sin(angles), np.cos(angles))
        elif any(t != 0 for t in angles):
            matrix = _rotation(angles[::-1])
    if shift is not None and any(s != 0 for s in shift):
        move = np.eye(nd + 1)
        move[:nd, nd] = shift[::-1]
        matrix = move if matrix is None else move @ matrix
    if zoom is not None and zoom != 1:
        scale = np.diag([zoom] * nd + [1.0])
        matrix = scale if matrix is None else scale @ matrix

    xform = None
    if matrix is not None:
        centre = np.asarray(spatial[::-1], dtype=np.float64) / 2.0 + 0.5
        fwd, back = np.eye(nd + 1), np.eye(nd + 1)
        fwd[:nd, nd] = centre
        back[:nd, nd] = -centre
        full = fwd @ matrix @ back
        xform = np.zeros((3, 4))
        xform[2, 2] = 1.0  # 2-D: identity z row (the kernel's D == 1 case)
        xform[:nd, :nd] = full[:nd, :nd]
        xform[:nd, 3] = full[:nd, nd]
    flags = 0
    if flip is not None:
        assert len(flip) == nd
        for i, f in enumerate(flip):
            if f and rng.random() < 0.5:
                flags |= 1 << (i + 3 - nd)
    return xform, flags


class ImageTransform:
    """Random affine augmentation (rotation, shift, zoom, flips; nearest neighbour; `cval` outside) of CUDA tensors.

    Args: as the reference's ImageTransform (dataset.py:63-92): rotation_range (degrees; scalar for 2-D, three entries
    (depth, height, width) for 3-D), shift_range (fractions of the size per axis), zoom_range (min, max), flip (one bool
    per axis), cval, augmentation_probability, seed.  None = not performed.

    `transform(x, y)` takes one sample, (C, D, H, W) or (C, H, W), like the reference's __call__; `batch(x, y)` takes
    (B, C, *spatial) and draws one set of parameters per sample in order -- sample b of batch n gets the draws the
    reference's transform would make on its (n * B + b)-th call.  x: float32 or int16; y: any integer type that holds
    the labels (uint8 is moved as it is, others through uint8 / int16).  Returns new tensors."""

    def __init__(self, rotation_range=None, shift_range=None, zoom_range=None, flip=None, cval=0., augmentation_probability=1.0,
                 seed=None):
        self.rotation_range = rotation_range
        self.shift_range = shift_range
        self.zoom_range = zoom_range
        self.flip = flip
        self.cval = cval
        self.augmentation_probability = augmentation_probability
        self.rng = np.random.default_rng(seed)

    def draw(self, spatial):
        return draw_transform(self.rng, tuple(spatial), self.rotation_range, self.shift_range, self.zoom_range, self.flip,
                              self.augmentation_probability)

    def __call__(self, x, y=None):
        out = self.batch(x[None], None if y is None else y[None])
        if y is None:
            return out[0]
        return out[0][0], out[1][0]

    transform = __call__

    def batch(self, x, y=None, params=None, out_x=None, out_y=None):
        """`params`: optional list of (xform, flags) per sample (from `draw`), e.g. to replay an augmentation.
        `out_x` / `out_y`: optional destination tensors (same shape and dtype as x / y; must not alias them), e.g. buffers whose
        addresses a captured CUDA graph holds; when given they are always written (a plain copy if nothing was drawn)."""
        spatial = tuple(x.shape[2:])
        if len(spatial) not in (2, 3):
            raise ValueError(f'ImageTransform.batch expects (B, C, H, W) or (B, C, D, H, W), got {tuple(x.shape)}')
        B = x.shape[0]
        if params is None:
            params = [self.draw(spatial) for _ in range(B)]
        if all(p[0] is None and p[1] == 0 for p in params) and out_x is None and out_y is None:
            return x if y is None else (x, y)  # the reference returns its inputs untouched, too
        host = np.zeros((B, 12))
        flags = np.zeros((B,), dtype=np.int32)
        for b, (xf, fl) in enumerate(params):
            if xf is None:
                flags[b] = fl | 8
            else:
                host[b] = np.asarray(xf, dtype=np.float64).reshape(12)
                flags[b] = fl
        dev = x.device
        xf_dev = torch.from_numpy(host).to(dev)
        fl_dev = torch.from_numpy(flags).to(dev)
        xo = _resample(x, xf_dev, fl_dev, self.cval, out_x)
        if y is None:
            return xo
        if tuple(y.shape[2:]) != spatial or y.shape[0] != B:
            raise ValueError('ImageTransform.batch: x and y must share the batch size and the spatial shape')
        return xo, _resample(y, xf_dev, fl_dev, self.cval, out_y)


_BYTES = {torch.uint8: 1, torch.int16: 2, torch.float32: 4}


def _resample(t, xf_dev, fl_dev, cval, out=None):
    if not isinstance(t, torch.Tensor) or t.device.type != 'cuda':
        raise RuntimeError('hno_b200: ImageTransform works on CUDA tensors; this package has no CPU path')
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f'hno_b200: tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}')
    orig = t.dtype
    if orig not in _BYTES:
        if orig.is_floating_point:
            t = t.to(torch.float32)
        elif orig in (torch.int8, torch.bool):
            t = t.to(torch.int16)
        else:  # int32 / int64 label maps: class indices fit 16 bits
            t = t.to(torch.int16)
    t = t.contiguous()
    B, C = t.shape[:2]
    sp = tuple(t.shape[2:])
    D, H, W = (1,) + sp if len(sp) == 2 else sp
    if out is not None:
        if (out.dtype != t.dtype or orig != t.dtype or out.shape != t.shape or not out.is_contiguous() or out.device != t.device
                or out.data_ptr() == t.data_ptr()):
            raise ValueError('ImageTransform: out must be a distinct contiguous tensor of the shape and dtype of its input '
                             '(uint8, int16 or float32)')
    else:
        out = torch.empty_like(t)
    call('hno_affine_resample_nn', ptr(t), ptr(out), _BYTES[t.dtype], ptr(xf_dev), ptr(fl_dev), B, C, D, H, W,
         float(cval), stream_ptr())
    return out if out.dtype == orig else out.to(orig)
