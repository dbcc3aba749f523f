"""ctypes binding of libsatb200.so (the C ABI declared in include/satb200.h).

The library is the only compute path of this package: if it is missing, or if a
tensor is not on a CUDA device, the calls below raise - there is no CPU or eager
PyTorch fallback (the oracle under ``oracle/`` is test infrastructure and is never
imported from here).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libsatb200.so")

_lib = None


class NativeError(RuntimeError):
    pass


class SatbDitConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "io_channels", "embed_dim", "depth", "num_heads", "cond_token_dim", "global_cond_dim",
        "project_cond_tokens", "project_global_cond", "global_cond_type", "patch_size", "operand_dtype", "qk_norm",
        "input_concat_dim", "prepend_cond_dim")]


SATB_MAX_STAGES = 8


class SatbOobleckConfig(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int), ("channels", ctypes.c_int), ("latent_dim", ctypes.c_int),
                ("n_stages", ctypes.c_int), ("c_mults", ctypes.c_int * SATB_MAX_STAGES),
                ("strides", ctypes.c_int * SATB_MAX_STAGES), ("final_tanh", ctypes.c_int),
                ("is_decoder", ctypes.c_int), ("operand_dtype", ctypes.c_int)]


# name -> (restype, argtypes); must list every symbol of include/satb200.h
_VP, _I, _LL, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
SIGNATURES = {
    "satb_last_error": (ctypes.c_char_p, []),
    "satb_abi_version": (_I, []),
    "satb_launch_count": (ctypes.c_ulonglong, []),
    "satb_reset_launch_count": (None, []),
    "satb_add_launch_count": (None, [ctypes.c_ulonglong]),
    "satb_dit_create": (_I, [ctypes.POINTER(SatbDitConfig), ctypes.POINTER(_VP)]),
    "satb_dit_destroy": (None, [_VP]),
    "satb_dit_load_weight": (_I, [_VP, ctypes.c_char_p, _VP, _LL, _VP]),
    "satb_dit_finalize": (_I, [_VP, _VP]),
    "satb_dit_reserve": (_I, [_VP, _I, _I]),
    "satb_dit_set_prepend_cond": (_I, [_VP, _VP, _I, _I, _VP]),
    "satb_dit_prepare_cond": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "satb_dit_forward": (_I, [_VP, _VP, _VP, _VP, _I, _I, _F, _F, _VP]),
    "satb_dit_forward_debug": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _F, _F, _VP]),
    "satb_dit_profile": (_I, [_VP, _I]),
    "satb_dit_profile_read": (_I, [_VP, _VP, _VP]),
    "satb_snake_beta": (_I, [_VP, _VP, _VP, _VP, _I, _I, _LL, _I, _VP]),
    "satb_sampler_update": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _LL] + [_F] * 8 + [_VP]),
    "satb_layernorm": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "satb_linear_f32out": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "satb_attention": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "satb_debug_attention_occupancy": (_I, [_I, _I]),
    "satb_attention_trace": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP, _VP]),
    "satb_oobleck_create": (_I, [ctypes.POINTER(SatbOobleckConfig), ctypes.POINTER(_VP)]),
    "satb_oobleck_destroy": (None, [_VP]),
    "satb_oobleck_load_weight": (_I, [_VP, ctypes.c_char_p, _VP, _LL, _VP]),
    "satb_oobleck_finalize": (_I, [_VP, _VP]),
    "satb_oobleck_decode": (_I, [_VP, _VP, _VP, _I, _I, _VP]),
    "satb_oobleck_encode": (_I, [_VP, _VP, _VP, _I, _LL, _VP]),
}


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python friendly-stable-audio-tools_b200/build.py` "
                "(or __graft_entry__.build()); this package has no non-CUDA fallback")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().satb_last_error()
        raise NativeError(f"satb200 error {rc}: {msg.decode() if msg else '?'}")


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dev_f32(t, name="tensor"):
    """Validate a CUDA fp32 contiguous tensor and return its device pointer."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise NativeError(f"{name} must be a CUDA tensor: this package runs on the GPU only (no CPU fallback)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise NativeError(f"{name} must be contiguous float32")
    return ctypes.c_void_p(t.data_ptr())


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def launch_count():
    return int(lib().satb_launch_count())
