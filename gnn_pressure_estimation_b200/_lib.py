"""ctypes binding of libgatres_b200.so (C ABI declared in include/gatres_b200.h).

This is the whole FFI: plain pointers, sizes and a stream handle — no torch
types cross the boundary.  There is NO fallback: if the shared library is
missing, importing the package's compute path raises with build instructions.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgatres_b200.so")
ABI_VERSION = 3

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float
_sz = C.c_size_t


class ModelDesc(C.Structure):
    """struct gatres_model_desc"""
    _fields_ = [("num_blocks", _i32), ("nc", _i32), ("N", _i32), ("slots", _i32), ("E1", _i32), ("reserved", _i32),
                ("B", _i64),
                ("rowptr", _p), ("col", _p), ("rowptr_t", _p), ("col_t", _p), ("poison", _p),
                ("perm", _p), ("p_rowptr", _p), ("p_col", _p), ("p_rowptr_t", _p), ("p_col_t", _p), ("p_ecap", _i32 * 4)]


# name -> (restype, argtypes); mirrors include/gatres_b200.h one to one
_PROTOTYPES = {
    "gatres_abi_version": (C.c_int, []),
    "gatres_last_error": (C.c_char_p, []),
    "gatres_sm_count": (C.c_int, []),
    "gatres_launch_count": (_i64, []),
    "gatres_set_tile_min_batch": (_i64, [_i64]),
    "gatres_set_resident_max_batch": (_i64, [_i64]),
    "gatres_set_resident_cluster": (_i32, [_i32]),
    "gatres_set_resident_threads": (_i32, [_i32]),
    "gatres_set_resident_profile": (None, [_p, _i32]),
    "gatres_set_tensor_core": (C.c_int, [C.c_int]),
    "gatres_set_resident_dsm": (C.c_int, [C.c_int]),
    "gatres_set_resident_tc": (C.c_int, [C.c_int]),
    "gatres_set_resident_barrier": (C.c_int, [C.c_int]),
    "gatres_csr_scratch_bytes": (_sz, [_i64, _i32]),
    "gatres_csr_build": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gatres_csr_build_mean": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gatres_check_replicated": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _p]),
    "gatres_linear_att_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "gatres_gat_agg_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p]),
    "gatres_gat_agg_bwd": (C.c_int, [_p] * 16 + [_i64, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _p]),
    "gatres_mean_res_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "gatres_mean_res_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "gatres_mean_res_bwd_e1": (C.c_int, [_p, _p, _p, _i32, _p, _p, _i64, _i32, _i32, _p]),
    "gatres_linear_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i64, _i64, _i32, _i32, _i32, _p]),
    "gatres_encoder_fwd": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p]),
    "gatres_encoder_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _i64, _i64, _i64, _i32, _p]),
    "gatres_decoder_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _p]),
    "gatres_decoder_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i64, _i64, _i64, _i32, _i32, _p]),
    "gatres_reduce_partials": (C.c_int, [_p, _i64, _i32, _i64, _i64, _p, _p]),
    "gatres_model_desc_bytes": (_sz, []),
    "gatres_param_count": (_i64, [_i32, _i32]),
    "gatres_saved_floats": (_i64, [C.POINTER(ModelDesc)]),
    "gatres_scratch_floats": (_i64, [C.POINTER(ModelDesc), _i32]),
    "gatres_forward": (C.c_int, [C.POINTER(ModelDesc), _p, _p, _p, _p, _p, _p]),
    "gatres_backward": (C.c_int, [C.POINTER(ModelDesc), _p, _p, _p, _p, _p, _p, _p, _p]),
    "gatres_backward_range": (C.c_int, [C.POINTER(ModelDesc), _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _p]),
    "gatres_param_offset_of_block": (_i64, [_i32, _i32, _i32]),
    "gatres_masked_mse": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "gatres_apply_mask": (C.c_int, [_p, _p, _p, _i64, _p]),
    "gatres_generate_mask": (C.c_int, [C.c_uint64, C.c_uint64, _p, _p, _i64, _i32, _i32, _p, _p]),
    "gatres_mask_key": (C.c_uint32, [C.c_uint64, C.c_uint64, C.c_uint64]),
    "gatres_metrics_scratch_doubles": (_i64, []),
    "gatres_masked_metrics": (C.c_int, [_p, _p, _p, _i64, _f32, _f32, _f32, _p, _p, _p]),
    "gatres_adam_step": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _p]),
    "gatres_adam_step_peer": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), _i32, _i32, _p, _p, _p, _p, _i64, _f32, _f32, _f32,
                                        _f32, _f32, _f32, _p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib: Optional[C.CDLL] = None


class GatresError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library once and attach prototypes.  No GPU needed for this."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the GATRes hot path has no CPU or PyTorch fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gnn_pressure_estimation_b200/csrc`).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    got = lib.gatres_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libgatres_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
    if lib.gatres_model_desc_bytes() != C.sizeof(ModelDesc):
        raise ImportError(f"struct gatres_model_desc is {lib.gatres_model_desc_bytes()} bytes in the library but "
                          f"{C.sizeof(ModelDesc)} in the binding; rebuild")
    _lib = lib
    return lib


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GatresError("libgatres_b200 kernels need CUDA tensors (no CPU path exists)")
    if not t.is_contiguous():
        raise GatresError("libgatres_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gatres_last_error().decode("utf-8", "replace")
        raise GatresError(f"{what} failed (rc={rc}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
