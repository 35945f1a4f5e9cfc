"""ctypes binding of libsgb200.so -- the C-ABI boundary declared in include/sgb200.h.

Everything the product computes goes through the functions bound here; there is no PyTorch
or CPU fallback.  If the library is missing the import of the product modules fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

SG_F32, SG_BF16, SG_F16 = 0, 1, 2
SG_ENGINE_SIMT, SG_ENGINE_TC = 0, 1
ABI_VERSION = 15
SG_ACT_NONE, SG_ACT_GELU, SG_ACT_RELU_POST = 0, 1, 2

_vp, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float


class IgemmArgs(C.Structure):
    """sg_igemm_args (include/sgb200.h)."""

    _fields_ = [
        ("a", _vp), ("w", _vp), ("bias", _vp), ("residual", _vp), ("out_f32", _vp), ("out_act", _vp),
        ("partials", _vp), ("a_lo", _vp), ("w_lo", _vp),
        ("rows", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("taps", C.c_int32), ("act", C.c_int32), ("engine", C.c_int32), ("act_dtype", C.c_int32),
        ("out_dtype", C.c_int32),
    ]


# name -> (restype, argtypes): must list every symbol include/sgb200.h declares (tests check this)
PROTOTYPES = {
    "sg_abi_version": (_i, []),
    "sg_last_error": (C.c_char_p, []),
    "sg_device_check": (_i, [_i]),
    "sg_set_device": (_i, [_i]),
    "sg_set_pdl": (_i, [_i]),
    "sg_time_embed": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sg_conv_in_partials": (_i, [_i]),
    "sg_conv_in": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "sg_igemm_partials": (_i, [_i, _i, _i, _i]),
    "sg_igemm": (_i, [C.POINTER(IgemmArgs), _vp]),
    "sg_gn_apply": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "sg_gn_apply_vcat": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "sg_maxpool2": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sg_upsample_cat": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sg_layernorm": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp]),
    "sg_ln_inproj": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _i, _vp]),
    "sg_attn_tail": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _i, _vp]),
    "sg_attn_tail_outc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp,
                               _i, _vp]),
    "sg_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sg_attn_prep_tf32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sg_attention_tf32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sg_split_tf32": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "sg_conv_out": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "sg_cfg_update": (_i, [_vp, _vp, _i, _i, _f, _vp, _i, _vp, _vp, _u64, _i64, _vp]),
    "sg_step_advance": (_i, [_vp, _vp]),
    "sg_philox_normal": (_i, [_vp, _i, _i, _u64, _i64, _i, _vp]),
    "sg_to_uint8": (_i, [_vp, _i64, _vp, _vp]),
    "sg_to_uint8_wrap": (_i, [_vp, _i64, _vp, _vp]),
    "sg_pack_weights": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "sg_noise_images": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _u64, _i64, _vp, _vp, _vp]),
    "sg_ema_update": (_i, [_vp, _vp, _i64, _f, _f, _vp]),
    "sg_mse_scratch_doubles": (_i, []),
    "sg_mse": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "sg_vq_quantize": (_i, [_vp, _i64, _vp, _i, _i, _vp, _vp, _vp]),
    "sg_dec_in_proj": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sg_tconv2_u8": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class SgError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen libsgb200.so (built in-tree by spectrogramgenai_b200.build) and bind the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise SgError(
            f"{path} is missing: build it with `python -m spectrogramgenai_b200.build` "
            "(or __graft_entry__.build()).  There is no fallback path."
        )
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.sg_abi_version() != ABI_VERSION:
        raise SgError(f"libsgb200.so ABI {lib.sg_abi_version()} != binding ABI {ABI_VERSION}: rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().sg_last_error()
        raise SgError(f"{what or 'libsgb200'} failed (status {rc}): {msg.decode() if msg else '?'}")


_checked_devices = set()


def require_b200(device: torch.device):
    """No fallback: the device must be a CUDA sm_100 part.  The check does not change the current device; callers that
    launch on a non-current device wrap the launches in `torch.cuda.device(device)` (engine.UNetPlan.run does)."""
    if device.type != "cuda":
        raise SgError(f"spectrogramgenai_b200 runs only on a B200 (sm_100a) CUDA device, got device '{device}'")
    lib = load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        check(lib.sg_device_check(idx), "sg_device_check")
        _checked_devices.add(idx)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    return {torch.float32: SG_F32, torch.bfloat16: SG_BF16, torch.float16: SG_F16}[dt]
