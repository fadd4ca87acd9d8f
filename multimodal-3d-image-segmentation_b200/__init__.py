"""hno_b200 — B200 (sm_100a) native spectral hot path of HNOSeg-XS behind the reference's module API.

Import as ``multimodal_3d_image_segmentation_b200`` (the directory name carries the upstream repository's
hyphens; ``multimodal_3d_image_segmentation_b200/__init__.py`` at the repository root is the import shim).

    from multimodal_3d_image_segmentation_b200 import nets
    model = nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14), device='cuda')

Everything numerical runs in hand-written CUDA kernels from ``libhno_b200.so`` (C ABI: include/hno_b200.h).
There is no CPU implementation and no fallback: on a machine without the library or without a B200 the
ops raise.
"""
from . import _lib  # noqa: F401
from ._lib import HnoError, load as load_library  # noqa: F401

__version__ = '0.1.0'


def __getattr__(name):
    # lazy: importing the package must not require torch.cuda
    if name in ('nets', 'ops', 'plan', 'engine', 'parallel', 'experiments'):
        import importlib
        return importlib.import_module(f'{__name__}.{name}')
    raise AttributeError(name)
