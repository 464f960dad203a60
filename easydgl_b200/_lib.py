"""ctypes binding of libeasydgl_b200.so (include/easydgl_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception
is raised.  Nothing in this package computes the forward pass on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libeasydgl_b200.so")

EDGL_MODEL_EASYDGL = 0
EDGL_MODEL_CTSMA = 1


class EdglConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("max_batch", C.c_int32), ("seq_len", C.c_int32), ("num_units", C.c_int32),
        ("num_heads", C.c_int32), ("num_blocks", C.c_int32), ("num_events", C.c_int32), ("num_rows", C.c_int32),
        ("mark_rows", C.c_int32), ("topk", C.c_int32), ("time_scale", C.c_float), ("mask_id", C.c_int64),
        ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int
_SIGS = {
    "edgl_last_error": (C.c_char_p, []),
    "edgl_version": (_I, []),
    "edgl_launch_count": (C.c_int64, []),
    "edgl_crc32c": (C.c_uint32, [C.c_void_p, C.c_size_t, C.c_uint32]),
    "edgl_num_stages": (_I, []),
    "edgl_stage_name": (C.c_char_p, [_I]),
    "edgl_profile": (_I, [_P, _I]),
    "edgl_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), _I]),
    "edgl_create": (_I, [C.POINTER(EdglConfig), C.POINTER(_P)]),
    "edgl_destroy": (_I, [_P]),
    "edgl_get_config": (_I, [_P, C.POINTER(EdglConfig)]),
    "edgl_set_tensor": (_I, [_P, C.c_char_p, _I, _P, C.c_int64]),
    "edgl_commit": (_I, [_P, _P]),
    "edgl_forward_logits": (_I, [_P, _P, _P, _I, _P, _P]),
    "edgl_forward_topk": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "edgl_forward_topk_host": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "edgl_forward_topk_host_submit": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "edgl_forward_topk_host_wait": (_I, [_P, _I]),
    "edgl_forward_train_logits": (_I, [_P, _P, _P, _I, _P, _I, _P, _P]),
    "edgl_forward_train_loss": (_I, [_P, _P, _P, _I, _P, _P, _I, C.c_float, C.c_float, _P, _P]),
    "edgl_encode": (_I, [_P, _P, _P, _I, _P, _P]),
    "edgl_encode_packed": (_I, [_P, _P, _P, _I, _P, C.c_int64, _P]),
    "edgl_logits_topk": (_I, [_P, _P, C.c_int64, _P, _I, C.c_int64, _I, C.c_int64, _P, _P, _P]),
    "edgl_xchg_alloc": (_I, [C.c_int64, C.POINTER(_P), _P]),
    "edgl_xchg_open": (_I, [_P, C.POINTER(_P)]),
    "edgl_xchg_close": (_I, [_P]),
    "edgl_xchg_free": (_I, [_P]),
    "edgl_xchg_put_rows": (_I, [_P, _P, C.c_int64, _P, _I, _P, _P, _I, _I, C.c_uint32, _P]),
    "edgl_xchg_wait": (_I, [_P, _I, C.c_uint32, _P]),
    "edgl_logits_topk_p2p": (_I, [_P, _P, C.c_int64, _P, _I, C.c_int64, _I, _I, _P, _P, _I, _I, C.c_uint32, _P]),
    "edgl_topk_merge": (_I, [_P, _P, _I, _I, _I, C.c_int64, C.c_int64, _P, _P, _P]),
    "edgl_time_sinusoid_code": (_I, [_P, _I, _I, _I, _P, _P]),
    "edgl_time_function_code": (_I, [_P, _P, _P, C.c_int64, _I, _P, _P]),
    "edgl_time_attention": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I,
                                 _P, _P]),
    "edgl_row_nonzero": (_I, [_P, C.c_int64, _I, _P, _P]),
    "edgl_layernorm_last": (_I, [_P, _P, _P, C.c_int64, _I, C.c_float, _P, _P]),
    "edgl_embedding_lookup": (_I, [_P, _I, _I, _I, _I, _P, C.c_int64, _P, _P]),
    "edgl_embed": (_I, [_P, _P, _P, _I, _P, _P, _P, _P]),
    "edgl_attention_layer": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P]),
    "edgl_intensity": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P]),
    "edgl_layernorm": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "edgl_dense": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "edgl_dense_nk": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "edgl_dense_nk_f16": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "edgl_topk": (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _P]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libeasydgl_b200.so is missing (%s). Build it with `python -m easydgl_b200.build` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EdglError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().edgl_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("easydgl_b200: " + msg)
        raise EdglError("easydgl_b200 (code %d): %s" % (rc, msg))
