"""Numeric-type tags shared by the host layer, the C-ABI and serialized `metadata`.

The integer values are a wire format: they are stored in `state_dict()["metadata"]`
and passed through `meta_args`, so they match the reference enum one-for-one
(reference: gemlite/dtypes.py:8-29).  Everything else in this file (helpers,
tables) is specific to the B200 build -- there is no Triton here.
"""
from enum import Enum

import torch


class DType(Enum):
    FP32 = 0
    FP16 = 1
    BF16 = 2
    FP8 = 3
    FP8e4 = 3  # alias of FP8 (e4m3fn)
    INT8 = 4
    UINT8 = 5
    INT32 = 6
    UINT32 = 7
    FP8e5 = 8
    INT16 = 9
    UINT16 = 10
    INT64 = 11
    FP8e4nuz = 12
    FP8e5nuz = 13
    MXFP16 = 14
    MXBF16 = 15
    MXFP8 = 16
    MXFP4 = 17
    NVFP4 = 18
    E8M0 = 19


_TORCH_OF = {
    DType.FP32: torch.float32,
    DType.FP16: torch.float16,
    DType.BF16: torch.bfloat16,
    DType.FP8: torch.float8_e4m3fn,
    DType.INT8: torch.int8,
    DType.UINT8: torch.uint8,
    DType.INT32: torch.int32,
    DType.UINT32: torch.uint32,
    DType.FP8e5: torch.float8_e5m2,
    DType.INT16: torch.int16,
    DType.UINT16: torch.uint16,
    DType.INT64: torch.int64,
    DType.FP8e4nuz: torch.float8_e4m3fnuz,
    DType.FP8e5nuz: torch.float8_e5m2fnuz,
    # MX tags describe how activations enter the kernel
    DType.MXFP16: torch.float16,
    DType.MXBF16: torch.bfloat16,
    DType.MXFP8: torch.float8_e4m3fn,
    DType.MXFP4: torch.uint8,
    DType.NVFP4: torch.uint8,
    DType.E8M0: torch.float8_e8m0fnu,
}

# int value -> torch dtype (the form `meta_args` consumers index with)
DTYPE_TO_TORCH = {d.value: t for d, t in _TORCH_OF.items()}

TORCH_TO_DTYPE = {
    torch.float32: DType.FP32,
    torch.float16: DType.FP16,
    torch.bfloat16: DType.BF16,
    torch.int8: DType.INT8,
    torch.uint8: DType.UINT8,
    torch.int32: DType.INT32,
    torch.uint32: DType.UINT32,
    torch.int16: DType.INT16,
    torch.uint16: DType.UINT16,
    torch.int64: DType.INT64,
    torch.float8_e4m3fn: DType.FP8,
    torch.float8_e5m2: DType.FP8e5,
    torch.float8_e4m3fnuz: DType.FP8e4nuz,
    torch.float8_e5m2fnuz: DType.FP8e5nuz,
    torch.float8_e8m0fnu: DType.E8M0,
}

# word type used to hold `packing_bitwidth` bits of packed codes
PACKING_BITWIDTH_TO_TORCH_DTYPE = {
    8: torch.uint8,
    16: torch.int16,
    32: torch.int32,
    64: torch.int64,
}

FP8_DTYPES = [DType.FP8, DType.FP8e4, DType.FP8e5, DType.FP8e4nuz, DType.FP8e5nuz]
FP8_INT8_DTYPES = [DType.INT8] + FP8_DTYPES
MX_DTYPES = [DType.MXFP16, DType.MXBF16, DType.MXFP8, DType.MXFP4, DType.NVFP4]
MX_DTYPES_val = [d.value for d in MX_DTYPES]


def is_mx_dtype(input_dtype) -> bool:
    """True for the micro-scaled activation tags (accepts DType or its int value)."""
    if isinstance(input_dtype, DType):
        return input_dtype in MX_DTYPES
    return int(input_dtype) in MX_DTYPES_val


def to_dtype(x) -> DType:
    return x if isinstance(x, DType) else DType(int(x))


__all__ = [
    "DType", "DTYPE_TO_TORCH", "TORCH_TO_DTYPE", "PACKING_BITWIDTH_TO_TORCH_DTYPE",
    "FP8_DTYPES", "FP8_INT8_DTYPES", "MX_DTYPES", "MX_DTYPES_val", "is_mx_dtype", "to_dtype",
]
