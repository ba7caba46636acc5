"""ctypes binding of include/molnextr_b200.h.  The shared library is mandatory: importing this
module on a machine where it has not been built raises -- there is no Python/CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MNX_LIB_PATH selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("MNX_LIB_PATH") or os.path.join(_HERE, "lib", "libmolnextr_b200.so")

MNX_OK = 0
ENCODER_NONE, ENCODER_SWIN_B, ENCODER_CONVNEXT_B = 0, 1, 2

EXPORTS = [
    "mnx_create", "mnx_destroy", "mnx_last_error", "mnx_load_tensor", "mnx_finalize_weights",
    "mnx_preprocess", "mnx_encode", "mnx_decode_greedy", "mnx_decode_greedy_labels", "mnx_decode_beam", "mnx_atom_indices", "mnx_edges", "mnx_confidence", "mnx_predict",
    "mnx_predict_host", "mnx_set_encoder_cta_limit", "mnx_beam_trace", "mnx_launch_count", "mnx_last_decode_steps", "mnx_time_kernel",
    "mnx_test_gemm_bf16", "mnx_reserve_contexts", "mnx_set_context", "mnx_set_decode_path", "mnx_set_wide_rows",
]


class MnxConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("encoder_kind", C.c_int32), ("max_batch", C.c_int32),
        ("max_height", C.c_int32), ("max_width", C.c_int32), ("max_len", C.c_int32),
        ("vocab", C.c_int32), ("tok_offset", C.c_int32), ("max_x", C.c_int32), ("max_y", C.c_int32),
        ("max_atoms", C.c_int32), ("encoder_dim", C.c_int32), ("max_beam", C.c_int32),
        ("token_class", C.POINTER(C.c_uint8)),
    ]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m molnextr_b200.build` "
            "(or __graft_entry__.build()).  molnextr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.mnx_create.argtypes = [C.POINTER(MnxConfig), C.POINTER(vp)]
    lib.mnx_destroy.argtypes = [vp]
    lib.mnx_last_error.argtypes = [vp]
    lib.mnx_last_error.restype = C.c_char_p
    lib.mnx_load_tensor.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32, i32]
    lib.mnx_finalize_weights.argtypes = [vp]
    lib.mnx_preprocess.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.mnx_encode.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.mnx_decode_greedy.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.mnx_decode_greedy_labels.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp, vp, vp, vp]
    lib.mnx_decode_beam.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.mnx_atom_indices.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.mnx_edges.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
    lib.mnx_confidence.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.mnx_predict.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.mnx_predict_host.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.mnx_set_encoder_cta_limit.argtypes = [vp, i32]
    lib.mnx_reserve_contexts.argtypes = [vp, i32]
    lib.mnx_set_context.argtypes = [vp, i32]
    lib.mnx_set_decode_path.argtypes = [vp, i32]
    lib.mnx_set_wide_rows.argtypes = [vp, i32]
    lib.mnx_beam_trace.argtypes = [vp, vp, i32]
    lib.mnx_launch_count.argtypes = [vp]
    lib.mnx_launch_count.restype = i64
    lib.mnx_last_decode_steps.argtypes = [vp]
    lib.mnx_last_decode_steps.restype = i32
    lib.mnx_time_kernel.argtypes = [vp, i32, i32, C.POINTER(C.c_float), vp]
    lib.mnx_test_gemm_bf16.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:   # default
            fn.restype = i32
    _lib = lib
    return lib
