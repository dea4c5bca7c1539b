"""ctypes binding of libcgpt_b200.so (the C ABI declared in include/cgpt.h).

There is no fallback: if the shared library is missing or the device is not a B200 (sm_100),
every compute entry point raises.  Loading the library itself needs no GPU, so CPU-only test
boxes can still check the exported symbols.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgpt_b200.so")

EPI_NONE, EPI_GELU, EPI_MUL_AUX = 0, 1, 2

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [("a", _vp), ("b", _vp), ("a_mn_major", _i), ("b_mn_major", _i), ("lda", _i64), ("ldb", _i64),
                ("M", _i), ("N", _i), ("K", _i), ("split_k", _i), ("bias", _vp), ("epilogue", _i), ("aux", _vp),
                ("aux_out", _vp), ("ldaux", _i64), ("residual", _vp), ("out", _vp), ("out_f32", _i),
                ("accumulate", _i), ("ldc", _i64), ("colsum", _vp)]


# name -> (restype, argtypes); must list every symbol include/cgpt.h declares (tests check this)
SIGNATURES = {
    "cgpt_version": (_i, []),
    "cgpt_device_ok": (_i, [_i]),
    "cgpt_set_device": (_i, [_i]),
    "cgpt_last_error": (C.c_char_p, []),
    "cgpt_launch_count": (_i64, []),
    "cgpt_segment_ids": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "cgpt_segment_starts": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "cgpt_next_in_set": (_i, [_vp, _vp, _i, _i, C.POINTER(_i64), _i, _vp]),
    "cgpt_termination_labels": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(_i64), _i, _i64, _vp]),
    "cgpt_embed_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cgpt_embed_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cgpt_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "cgpt_layernorm_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cgpt_gemm_bf16": (_i, [C.POINTER(GemmArgs), _vp]),
    "cgpt_cast_f32_bf16": (_i, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "cgpt_split3_f32_bf16": (_i, [_vp, _i64, _vp, _i64, _i64, _i64, _i, _vp]),
    "cgpt_fold_quadrants_add": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp]),
    "cgpt_colsum_bf16": (_i, [_vp, _i64, _vp, _i, _i, _vp]),
    "cgpt_rope_qk": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "cgpt_swiglu_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp]),
    "cgpt_swiglu_bwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _vp]),
    "cgpt_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, C.c_uint64, C.c_uint64, _vp]),
    "cgpt_attn_bwd_workspace": (_i64, [_i, _i, _i, _i, _i]),
    "cgpt_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, C.c_uint64, C.c_uint64,
                           _vp]),
    "cgpt_attn_bwd_colsum": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, C.c_uint64,
                                  C.c_uint64, _vp]),
    "cgpt_attn_probs": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, C.c_uint64, C.c_uint64, _vp]),
    "cgpt_attn_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "cgpt_pack_lm_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "cgpt_shape_proj_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "cgpt_shape_proj_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "cgpt_dropout": (_i, [_vp, _vp, _vp, _i, _i64, _f, C.c_uint64, C.c_uint64, _vp]),
    "cgpt_skinny_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cgpt_skinny_linear_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "cgpt_ce_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i64, _i, _vp]),
    "cgpt_ce_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i64, _i, _i, _i, _i, _f, _i64, _vp]),
    "cgpt_set_philox_state": (_i, [_vp]),
    "cgpt_philox_advance": (_i, [_vp, C.c_uint64, _vp]),
    "cgpt_adamw": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _vp]),
    "cgpt_adamw_bf16grad": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _vp]),
}


class CgptError(RuntimeError):
    pass


_lib = None
_device_checked = set()


def load():
    """Load the shared library (no GPU needed).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CgptError(
                f"{LIB_PATH} is missing: build it with `python genomics-lm_b200/build.py` "
                "(there is no CPU / PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().cgpt_last_error()
        raise CgptError(f"cgpt error {rc}: {msg.decode() if msg else '?'}")


def ensure_device(index: int):
    """Fail loudly unless `index` is a B200-class device; bind this thread to it."""
    if index in _device_checked:
        return
    lib = load()
    check(lib.cgpt_set_device(index))
    check(lib.cgpt_device_ok(index))
    _device_checked.add(index)
