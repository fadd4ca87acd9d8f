"""`to_categorical` and `normalize_modalities` of the reference's experiments/utils.py on CUDA tensors.

Same names, argument meaning and results as experiments/utils.py:25-97; the inputs are CUDA tensors (the reference's
normalize_modalities works on numpy arrays inside the loader workers, its to_categorical on whatever device the label
batch is on).  No CPU path: a non-CUDA tensor raises.
"""
import torch

from .. import _lib
from .._lib import call, ptr, stream_ptr
from ..ops import workspace


def _cuda(t, name):
    if not isinstance(t, torch.Tensor) or t.device.type != 'cuda':
        raise RuntimeError(f'hno_b200: {name} must be a CUDA tensor; this package has no CPU path')


def to_categorical(y, num_classes=None, validate=True):
    """Integer labels (B, 1, *spatial) -> one-hot float32 (B, num_classes, *spatial)  (experiments/utils.py:74-97).

    uint8 and int64 labels are read directly; other integer types are converted to int64 first (the reference converts
    everything with `.to(dtype=int)`).  `num_classes=None` uses `y.max() + 1` like the reference (a device sync).  With
    `validate` (default) a label outside [0, num_classes) raises IndexError, as the reference's scatter does; this reads
    one int back from the device -- pass validate=False inside a CUDA graph or to avoid the sync.
    """
    _cuda(y, 'y')
    assert y.shape[1] == 1, 'Can only handle single label per pixel.'
    if y.dtype not in (torch.uint8, torch.int64):
        if y.dtype.is_floating_point or y.dtype in (torch.int8, torch.int16, torch.int32, torch.bool):
            y = y.to(torch.int64)
        else:
            raise TypeError(f'to_categorical: unsupported label dtype {y.dtype}')
    y = y.contiguous()
    if not num_classes:
        num_classes = int(y.max()) + 1
    num_classes = int(num_classes)
    B = y.shape[0]
    spatial = tuple(y.shape[2:])
    N = 1
    for s in spatial:
        N *= s
    out = torch.empty((B, num_classes) + spatial, dtype=torch.float32, device=y.device)
    bad = torch.empty((1,), dtype=torch.int32, device=y.device) if validate else None
    call('hno_to_categorical', ptr(y), y.element_size(), ptr(out), ptr(bad), B, num_classes, N, stream_ptr())
    if validate and int(bad) != 0:
        raise IndexError(f'to_categorical: {int(bad)} labels are outside [0, {num_classes})')
    return out


def normalize_modalities(data, mask_val=None, clip_val=None):
    """Normalises every slice along the first axis (one modality each) separately  (experiments/utils.py:25-71):
    optional clip to `clip_val = (min, max)`, mean / std over the voxels that differ from `mask_val` (after clipping; all
    voxels when mask_val is None), (x - mean) / std, masked voxels set to 0.  Returns float32.

    int16 input (the storage type of the NIfTI volumes) is read as it is -- the conversion to float32 is exact and
    happens inside the kernels, so a raw batch can cross PCIe at half the bytes; `out` lets the caller pass the
    destination (e.g. a buffer whose address a captured CUDA graph already holds)."""
    return normalize_rows(data, data.shape[0], mask_val, clip_val)


def normalize_rows(data, rows, mask_val=None, clip_val=None, out=None):
    """normalize_modalities with the number of independently normalised rows given explicitly: a batch (B, C, *spatial)
    is B * C rows (each sample and modality separately, which is what the reference's per-sample x_processing does)."""
    _cuda(data, 'data')
    x = data.contiguous() if data.dtype in (torch.float32, torch.int16) else data.to(torch.float32).contiguous()
    rows = int(rows)
    n = x.numel() // rows
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    elif out.dtype != torch.float32 or out.numel() != x.numel() or not out.is_contiguous() or out.device != x.device:
        raise ValueError('normalize_rows: out must be a contiguous float32 CUDA tensor of the size of data')
    ws = workspace(_lib.load().hno_normalize_workspace_bytes(rows), x.device, 'normalize')
    lo, hi = (float(clip_val[0]), float(clip_val[1])) if clip_val is not None else (0.0, 0.0)
    fn = 'hno_normalize_modalities_i16' if x.dtype == torch.int16 else 'hno_normalize_modalities'
    call(fn, ptr(x), ptr(out), ptr(ws), rows, n, int(mask_val is not None),
         float(mask_val) if mask_val is not None else 0.0, int(clip_val is not None), lo, hi, stream_ptr())
    return out
