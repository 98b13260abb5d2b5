"""Helpers to read tests/golden/*.npz (raw-bit fixtures written by make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN_DIR, "MANIFEST.json")))

_T = {
    "torch.float16": torch.float16, "torch.bfloat16": torch.bfloat16, "torch.float32": torch.float32,
    "torch.int8": torch.int8, "torch.uint8": torch.uint8, "torch.int32": torch.int32,
    "torch.int16": torch.int16, "torch.int64": torch.int64,
    "torch.float8_e4m3fn": torch.float8_e4m3fn, "torch.float8_e5m2": torch.float8_e5m2,
    "torch.float8_e8m0fnu": torch.float8_e8m0fnu,
}


def tdtype(name):
    return _T[name]


def from_bits(arr, dtype):
    """numpy raw bits -> torch tensor of `dtype` (exact)."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dtype in (torch.float16, torch.bfloat16):
        return t.view(torch.int16).view(dtype) if t.dtype != torch.uint16 else t.view(dtype)
    if dtype in (torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e8m0fnu):
        return t.view(dtype)
    return t.to(dtype) if t.dtype != dtype else t


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz")), MANIFEST[name]


NOT_LAYERS = ("bitpack", "actq_mx", "weight_quant_mx")          # fixtures that are not a packed layer + forward outputs
# MXFP / NVFP weights with dynamically quantised activations
MX_DYNAMIC = {"a8w8_mxfp_dynamic_post": ("a8w8", True), "a8w8_mxfp_dynamic": ("a8w8", False),
              "a8w4_mxfp_dynamic": ("a8w4", False), "a8w4_mxfp_dynamic_post": ("a8w4", True),
              "a4w4_mxfp_dynamic": ("a4w4_mx", False), "a4w4_nvfp_dynamic": ("a4w4_nv", False)}


def layer_cases():
    return sorted(k for k, v in MANIFEST.items() if not k.startswith("_") and k not in NOT_LAYERS)
