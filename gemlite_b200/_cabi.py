"""ctypes binding of libgemlite_b200.so (the C ABI declared in include/gemlite_b200.h).

The shared object is built in-tree by `gemlite_b200.build` (nvcc, sm_100a).  It is loaded lazily;
if it is missing or fails to load, every op raises -- there is no CPU or PyTorch fallback.
"""
import ctypes
import os
import threading

import torch

from .dtypes import TORCH_TO_DTYPE, DType

_HERE = os.path.dirname(os.path.abspath(__file__))
# GLB_LIB_PATH: diagnostics only (e.g. the phase-timeline build, `python -m gemlite_b200.build --timeline`)
LIB_PATH = os.environ.get("GLB_LIB_PATH") or os.path.join(_HERE, "libgemlite_b200.so")

_lib = None
_lock = threading.Lock()


class GlbLinear(ctypes.Structure):
    """Mirror of `struct glb_linear` (field order and types must match the header)."""
    _fields_ = [
        ("W_q", ctypes.c_void_p), ("wq_rows", ctypes.c_int64), ("wq_cols", ctypes.c_int64),
        ("wq_stride0", ctypes.c_int64), ("wq_stride1", ctypes.c_int64), ("wq_dtype", ctypes.c_int32),
        ("scales", ctypes.c_void_p), ("scales_numel", ctypes.c_int64),
        ("s_stride0", ctypes.c_int64), ("s_stride1", ctypes.c_int64), ("scales_dtype", ctypes.c_int32),
        ("zeros", ctypes.c_void_p), ("zeros_numel", ctypes.c_int64),
        ("z_stride0", ctypes.c_int64), ("z_stride1", ctypes.c_int64), ("zeros_dtype", ctypes.c_int32),
        ("_pad0", ctypes.c_int32),
        ("K", ctypes.c_int64), ("N", ctypes.c_int64),
        ("scaled_activations", ctypes.c_int32), ("W_nbits", ctypes.c_int32), ("group_size", ctypes.c_int32),
        ("unpack_mask", ctypes.c_int32), ("elements_per_sample", ctypes.c_int32), ("input_dtype", ctypes.c_int32),
        ("output_dtype", ctypes.c_int32), ("acc_dtype", ctypes.c_int32), ("meta_dtype", ctypes.c_int32),
        ("channel_scale_mode", ctypes.c_int32), ("W_group_mode", ctypes.c_int32), ("data_contiguous", ctypes.c_int32),
    ]


EXPORTS = [
    "glb_version", "glb_last_error", "glb_last_kernel", "glb_launch_count", "glb_pack_cols", "glb_unpack_cols",
    "glb_quant_act_per_token", "glb_quant_act_mx", "glb_stream_capture_id", "glb_workspace_bytes", "glb_matmul", "glb_set_pdl", "glb_get_pdl", "glb_debug_timeline", "glb_matmul_bcast", 
]


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises NativeLibraryError if the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found. Build it with `python -m gemlite_b200.build` "
                "(gemlite_b200 has no CPU / PyTorch fallback).")
        lib = ctypes.CDLL(LIB_PATH)
        vp, i64, i32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
        lib.glb_version.restype = i32
        lib.glb_last_error.restype = ctypes.c_char_p
        lib.glb_last_kernel.restype = ctypes.c_char_p
        lib.glb_launch_count.restype = ctypes.c_uint64
        lib.glb_pack_cols.argtypes = [vp, i64, i64, i32, i32, vp, i32, vp]
        lib.glb_pack_cols.restype = i32
        lib.glb_unpack_cols.argtypes = [vp, i64, i64, i32, i32, vp, i32, vp]
        lib.glb_unpack_cols.restype = i32
        lib.glb_quant_act_per_token.argtypes = [vp, i64, i64, i32, i32, vp, vp, vp]
        lib.glb_quant_act_per_token.restype = i32
        lib.glb_quant_act_mx.argtypes = [vp, i64, i64, i32, i32, vp, vp, vp]
        lib.glb_quant_act_mx.restype = i32
        lib.glb_stream_capture_id.argtypes = [vp]
        lib.glb_stream_capture_id.restype = ctypes.c_ulonglong
        lib.glb_workspace_bytes.argtypes = [ctypes.POINTER(GlbLinear), i64, i32]
        lib.glb_workspace_bytes.restype = sz
        lib.glb_matmul.argtypes = [ctypes.POINTER(GlbLinear), vp, vp, vp, i64, i32, vp, sz, vp, vp]
        lib.glb_matmul.restype = i32
        lib.glb_matmul_bcast.argtypes = [ctypes.POINTER(GlbLinear), vp, vp, ctypes.POINTER(ctypes.c_void_p), i32, i64, i64, vp, vp]
        lib.glb_matmul_bcast.restype = i32
        lib.glb_set_pdl.argtypes = [i32]
        lib.glb_set_pdl.restype = None
        lib.glb_get_pdl.restype = i32
        lib.glb_debug_timeline.argtypes = [vp, i32]
        lib.glb_debug_timeline.restype = i32
        if os.environ.get("GLB_PDL", "1") == "0":
            lib.glb_set_pdl(0)
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().glb_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def current_stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (gemlite_b200 runs on sm_100a only, no CPU fallback)")


def dtype_tag(t: torch.dtype) -> int:
    return TORCH_TO_DTYPE[t].value


def last_kernel() -> str:
    return load().glb_last_kernel().decode()


def launch_count() -> int:
    return int(load().glb_launch_count())

