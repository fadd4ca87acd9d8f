"""ctypes binding of libhno_b200.so (the C ABI in include/hno_b200.h).

There is deliberately no fallback: if the shared object is missing it is built with nvcc, and if that
fails (or a call returns an error) an exception is raised.  Nothing in this package computes on the CPU.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libhno_b200.so')


class HnoError(RuntimeError):
    """An hno_b200 entry point reported failure."""


_P = c_void_p
_I = c_int
_L = c_long
_F = c_float
_Z = c_size_t
_D = c_double

# name: (restype, argtypes)  -- mirrors include/hno_b200.h one to one
SIGNATURES = {
    'hno_version': (_I, []),
    'hno_last_error': (c_char_p, []),
    'hno_device_check': (_I, []),
    'hno_launch_count': (_L, [_I]),
    'hno_set_tensor_cores': (_I, [_I]),
    'hno_dht3_plan_bytes': (_Z, [_I] * 6),
    'hno_dht3_plan_fill': (_I, [_P, _Z, _I, _I, _I, _P, _I, _P, _I, _P, _I]),
    'hno_dht3_workspace_bytes': (_Z, [_P, _L, _I]),
    'hno_dht3_forward': (_I, [_P, _P, _P, _L, _L, _P, _P, _I, _F, _P]),
    'hno_dht3_adjoint': (_I, [_P, _P, _P, _P, _L, _L, _P, _I, _F, _I, _P]),
    'hno_dht3_chain_eligible': (_I, [_P, _P, _L, _L, _I, _I, _I]),
    'hno_dht3_chain_partials_bytes': (_Z, [_P, _I, _I, _I]),
    'hno_dht3_chain_forward': (_I, [_P, _P, _P, _P, _L, _L, _P, _P, _P, _I, _I, _I, _F, _I, _P]),
    'hno_dht3_chain_backward': (_I, [_P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _I, _P]),
    'hno_pwconv_supported': (_I, [_I, _I, _I]),
    'hno_pwconv_forward': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _I, _I, _P]),
    'hno_pwconv_backward_workspace_bytes': (_Z, [_I, _I, _I]),
    'hno_pwconv_backward': (_I, [_P] * 10 + [_I, _I, _I, _I, _L, _L, _L, _I, _I, _I, _P]),
    'hno_modechain_supported': (_I, [_I]),
    'hno_modechain_forward': (_I, [_P, _P, _P, _I, _I, _L, _I, _P]),
    'hno_modechain_backward_workspace_bytes': (_Z, [_I, _I, _L, _I]),
    'hno_modechain_backward': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _I, _P]),
    'hno_hartley_conv_forward': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'hno_hartley_conv_backward': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'hno_hartley_conv_full_forward': (_I, [_P, _P, _P, _P] + [_I] * 10 + [_P]),
    'hno_hartley_conv_full_backward': (_I, [_P] * 7 + [_I] * 9 + [_P]),
    'hno_complex_modemix_forward': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _L, _P]),
    'hno_complex_modemix_backward': (_I, [_P] * 10 + [_I, _I, _I, _L, _I, _P]),
    'hno_fourier_mix_workspace_bytes': (_Z, [_I, _I, _L, _I]),
    'hno_fourier_mix_forward': (_I, [_P] * 7 + [_I, _I, _I, _L, _L, _I, _P]),
    'hno_fourier_mix_backward': (_I, [_P] * 11 + [_I, _I, _I, _L, _L, _I, _I, _P]),
    'hno_stem_supported': (_I, [_I, _I]),
    'hno_stem_forward': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _P]),
    'hno_stem_backward_workspace_bytes': (_Z, [_I, _I]),
    'hno_stem_backward': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _I, _P]),
    'hno_stem_backward_input': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _P]),
    'hno_interp_tables_bytes': (_Z, [_I] * 6),
    'hno_interp_tables_fill': (_I, [_P, _Z] + [_I] * 6),
    'hno_head_forward': (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _P]),
    'hno_head_argmax': (_I, [_P, _P, _P, _P, _I, _I, _L, _P]),
    'hno_head_direct_forward': (_I, [_P, _P, _I, _I, _I, _I, _I, _L, _I, _P]),
    'hno_head_direct_argmax': (_I, [_P, _P, _I, _I, _I, _I, _I, _L, _P]),
    'hno_head_direct_backward': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _L, _I, _P]),
    'hno_head_backward_workspace_bytes': (_Z, [_P, _I, _I]),
    'hno_head_backward': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P]),
    'hno_loss_workspace_bytes': (_Z, [_I, _I]),
    'hno_loss_forward': (_I, [_P, _P, _P, _P, _P, _I, _I, _L, _I, _F, _P]),
    'hno_loss_backward': (_I, [_P, _P, _P, _P, _P, _I, _I, _L, _P]),
    'hno_head_loss_forward': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _F, _P]),
    'hno_head_loss_backward': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _P]),
    'hno_ce_loss_workspace_bytes': (_Z, [_I]),
    'hno_ce_loss_forward': (_I, [_P, _P, _P, _P, _P, _I, _I, _L, _P]),
    'hno_ce_loss_backward': (_I, [_P, _P, _P, _P, _P, _I, _I, _L, _P]),
    'hno_to_categorical': (_I, [_P, _I, _P, _P, _I, _I, _L, _P]),
    'hno_normalize_workspace_bytes': (_Z, [_I]),
    'hno_normalize_modalities': (_I, [_P, _P, _P, _I, _L, _I, _F, _I, _F, _F, _P]),
    'hno_normalize_modalities_i16': (_I, [_P, _P, _P, _I, _L, _I, _F, _I, _F, _F, _P]),
    'hno_affine_resample_nn': (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _D, _P]),
    'hno_transpose2d': (_I, [_P, _P, _I, _L, _I, _I, _P]),
    'hno_dsconv_forward': (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P]),
    'hno_dsconv_backward_workspace_bytes': (_Z, [_I, _I, _I, _L]),
    'hno_dsconv_backward': (_I, [_P, _P, _P, _I] + [_P] * 8 + [_I, _I, _L, _L, _L, _I, _P]),
    'hno_mha_project_forward': (_I, [_P] * 5 + [_I] * 12 + [_P]),
    'hno_mha_wgrad_workspace_bytes': (_Z, [_I] * 5),
    'hno_mha_project_backward': (_I, [_P] * 7 + [_I] * 13 + [_P]),
    'hno_mha_attention_forward': (_I, [_P] * 6 + [_I, _I, _I, _I, _F, _I, _P]),
    'hno_mha_attention_backward': (_I, [_P] * 12 + [_I, _I, _I, _I, _F, _I, _P]),
    'hno_mha_output_forward': (_I, [_P] * 4 + [_I] * 12 + [_P]),
    'hno_mha_output_backward': (_I, [_P] * 8 + [_I] * 12 + [_P]),
    'hno_adamax_step': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P]),
}

_lib = None


def load(build_if_missing=True):
    """Returns the loaded ctypes.CDLL; builds it first when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and not build_if_missing:
        raise HnoError(f'{LIB_PATH} is missing; run python multimodal-3d-image-segmentation_b200/build.py')
    if build_if_missing:
        # always through build(): it is a digest comparison when the library is current, and it rebuilds a stale one
        # (sources newer than the .so) instead of loading it silently.  One process builds at a time (file lock, the
        # .so is renamed into place), so the ranks of a torchrun job cannot dlopen a half-written library.
        import importlib.util
        spec = importlib.util.spec_from_file_location('_hno_b200_build', os.path.join(HERE, 'build.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().hno_last_error().decode('utf-8', 'replace')


def call(name, *args):
    """Calls an int-returning entry point and raises HnoError(message) on a negative status."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise HnoError(f'{name} failed ({rc}): {last_error()}')


def ptr(t):
    """Device/host address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
