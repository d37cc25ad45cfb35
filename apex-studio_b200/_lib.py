"""ctypes loader for libapex_b200.so -- the C ABI declared in include/apex_b200.h.

The library is built in-tree by ``__graft_entry__.build()`` (``make -C apex-studio_b200/csrc``) so it
travels with the repository snapshot.  There is no fallback: if the shared object is missing every op
raises (the north star forbids CPU / library fallbacks).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# APEX_B200_LIB: an alternative build of the same library (kernel experiments, scripts/sessions/*); default: the in-tree build
LIB_PATH = os.environ.get("APEX_B200_LIB") or os.path.join(_HERE, "libapex_b200.so")

_i, _i64, _f, _p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

class WanResBlock(ctypes.Structure):
    _fields_ = [(n, _p) for n in ("norm1_gamma", "conv1_w", "conv1_b", "norm2_gamma", "conv2_w", "conv2_b", "shortcut_w", "shortcut_b")] + \
               [("cin", _i), ("cout", _i)]


class WanAttn(ctypes.Structure):
    _fields_ = [(n, _p) for n in ("norm_gamma", "to_qkv_w", "to_qkv_b", "proj_w", "proj_b")] + [("channels", _i)]


class WanUpsample(ctypes.Structure):
    _fields_ = [(n, _p) for n in ("resample_w", "resample_b", "time_conv_w", "time_conv_b")] + [("channels", _i), ("temporal", _i)]


class WanVaeWeights(ctypes.Structure):
    """B200WanVaeWeights of include/apex_b200.h"""
    _fields_ = [("post_quant_w", _p), ("post_quant_b", _p), ("conv_in_w", _p), ("conv_in_b", _p), ("mid_res", WanResBlock * 2),
                ("mid_attn", WanAttn), ("up_res", (WanResBlock * 3) * 4), ("up_samp", WanUpsample * 3), ("norm_out_gamma", _p),
                ("conv_out_w", _p), ("conv_out_b", _p), ("dims", _i * 5), ("z_pad", _i)]


# name -> (restype, argtypes); must list every symbol include/apex_b200.h declares.
SIGNATURES = {
    "b200_version": (_i, []),
    "b200_strerror": (ctypes.c_char_p, [_i]),
    "b200_attn_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i] + [_i64] * 12 + [_f, _p]),
    "b200_attn_fwd_prof": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i] + [_i64] * 12 + [_f, _p, _i, _p]),
    "b200_linear": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "b200_linear_normw": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i64, _i64, _i64, _p]),
    "b200_attn_fwd_qnorm": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i] + [_i64] * 12 + [_f, _p, _i, _i, _f, _p]),
    "b200_layernorm_modulate": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _i64, _i64, _f, _p]),
    "b200_rmsnorm_rope": (_i, [_p, _p, _p, _i, _i, _i, _i64, _f, _p]),
    "b200_rmsnorm_rope_batched": (_i, [_p, _p, _p, _i, _i, _i, _i64, _f, _i, _i64, _i64, _p]),
    "b200_gate_residual": (_i, [_p, _p, _p, _i, _i, _i64, _i64, _p]),
    "b200_cfg_combine": (_i, [_p, _p, _p, _f, _i64, _p]),
    "b200_rmsnorm_rope_scatter": (_i, [_p, _p, _p, _i, _i, _i, _i64, _f, _p, _i, _i64, _i, _p]),
    "b200_attn_fwd_scatter": (_i, [_p, _p, _p, _i, _i, _i, _i] + [_i64] * 6 + [_p, _i, _i, _i, _i64, _i64, _f, _p]),
    "b200_attn_fwd_scatter_joint": (_i, [_p, _p, _p, _i, _i, _i, _i] + [_i64] * 6 + [_p, _i, _i, _i, _i64, _i64, _i, _i, _f, _p]),
    "b200_conv3d_cl": (_i, [_p, _p, _p, _p, _p] + [_i] * 13 + [_p]),
    "b200_rmsnorm_silu_cl": (_i, [_p, _p, _p, _i64, _i, _i, _p]),
    "b200_upsample2x_cl": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "b200_softmax_rows": (_i, [_p, _p, _i, _i, _i64, _i64, _f, _p]),
    "b200_blend_tile": (_i, [_p, _p, _p, _p] + [_i] * 14 + [_p]),
    "b200_frames_to_uint8": (_i, [_p, _p, _i, _i, _i, _p]),
    "b200_adaln_zero_modulate": (_i, [_p, _p, _p, _p, _i, _i, _i64, _i64, _f, _p]),
    "b200_headnorm_rope": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i64, _f, _i, _p]),
    "b200_rmsnorm_rows": (_i, [_p, _p, _p, _i, _i, _i64, _i64, _f, _i, _p]),
    "b200_swiglu": (_i, [_p, _p, _i, _i, _i64, _i64, _p]),
    "b200_pad_norm_silu_cl": (_i, [_p, _p, _p] + [_i] * 8 + [_p]),
    "b200_conv3d_cl_padded": (_i, [_p, _p, _p, _p, _p] + [_i] * 10 + [_p]),
    "b200_conv3d_cl_norm_silu": (_i, [_p, _p, _p, _p, _p] + [_i] * 8 + [_p]),
    "b200_dcae_upsample_cl": (_i, [_p, _p, _p] + [_i] * 6 + [_p]),
    "b200_softmax_rows_block_causal": (_i, [_p, _p, _i, _i, _i64, _i64, _f, _i, _p]),
    "b200_blend_tile_noclamp": (_i, [_p, _p, _p, _p] + [_i] * 14 + [_p]),
    "b200_wan_vae_decode_workspace": (_i64, [ctypes.POINTER(WanVaeWeights), _i, _i, _i]),
    "b200_wan_vae_decode": (_i, [_p, ctypes.POINTER(WanVaeWeights), _p, _p, _i64, _i, _i, _i, _p]),
    "b200_ln_modulate": (_i, [_p, _p, _p, _p, _i, _i, _f, _p]),
    "b200_qkv_rmsnorm_rope": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i64, _f, _p]),
    "b200_mlp_gelu": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
}

B200_OK = 0
_VALUE_ERRORS = {-1, -2, -6}  # shape / alignment / argument -> ValueError (as functions.py:791-801 raises)

_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C apex-studio_b200/csrc). There is no CPU or library fallback for this path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code == B200_OK:
        return
    msg = load().b200_strerror(code).decode()
    if code in _VALUE_ERRORS:
        raise ValueError(f"{what}: {msg} (code {code})")
    raise RuntimeError(f"{what}: {msg} (code {code})")
