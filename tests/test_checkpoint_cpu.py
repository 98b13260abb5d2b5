"""SURVEY 8f item 2 (on-disk format): checkpoints written by the REFERENCE load into gemlite_b200 unchanged.

Every layer fixture in tests/golden/ holds what the reference's `state_dict()` would save for that layer -- the packed
`W_q` (with the shape / strides pack() gave it), `scales`, `zeros`, the 12-int `metadata` parameter exactly as the
reference registered it (core.py:502-517; for the helper processors that override the modes after pack() it is STALE,
helper.py:156-163,474-475 -- a reference quirk a loader has to reproduce, not repair) and `orig_shape`.  This test
rebuilds that state dict, loads it with `GemLiteLinear.load_state_dict` (core.py:301-333 semantics) on the CPU (loading
is host logic; running the layer is the GPU suite's job) and checks every attribute a forward reads.  It also pins the
save side: `state_dict()` returns the reference's six keys with the same tensors, through a torch.save round trip."""
import io

import numpy as np
import pytest
import torch

from gemlite_b200 import DType, GemLiteLinear
from gemlite_b200.dtypes import DTYPE_TO_TORCH
from golden_io import from_bits, layer_cases, load, tdtype


def reference_state_dict(z, info):
    wdt = tdtype(info["W_q_dtype"])
    shape, strides = info["W_q_shape"], [int(v) for v in z["packed_W_q_strides"]]
    flat = from_bits(z["packed_W_q"], wdt).reshape(-1)
    if strides == [shape[1], 1]:
        W_q = flat.view(shape)
    else:                                       # (K, N) view with strides (1, K) of an (N, K) row-major tensor
        assert strides == [1, shape[0]], strides
        W_q = flat.view(shape[1], shape[0]).t()
    sdt, zdt = tdtype(info["scales_dtype"]), tdtype(info["zeros_dtype"])
    scales = from_bits(z["packed_scales"], sdt).view(info["scales_shape"])
    zeros = from_bits(z["packed_zeros"], zdt).view(info["zeros_shape"])
    meta = torch.from_numpy(z["metadata_param"].astype(np.int32))
    K = shape[0] * int(meta[4])
    N = shape[1]
    sd = {"W_q": W_q, "scales": scales, "zeros": zeros, "metadata": meta,
          "orig_shape": torch.tensor([N, K], dtype=torch.int32)}
    if "in_bias" in z.files:
        sd["bias"] = from_bits(z["in_bias"], torch.float16)
    return sd, N, K


@pytest.mark.parametrize("name", layer_cases())
def test_reference_checkpoint_loads(name):
    z, info = load(name)
    sd, N, K = reference_state_dict(z, info)
    want = dict(sd)
    layer = GemLiteLinear.__new__(GemLiteLinear)
    torch.nn.Module.__init__(layer)
    layer.load_state_dict(dict(sd))
    meta = [int(v) for v in z["metadata_param"]]
    assert layer.get_meta_args() == meta                         # the (possibly stale) saved modes, as the reference would
    assert (layer.out_features, layer.in_features) == (N, K)
    assert layer.W_q.shape[0] * layer.elements_per_sample == K and layer.W_q.shape[1] == N
    assert list(layer.W_q.stride()) == [int(v) for v in z["packed_W_q_strides"]]
    assert isinstance(layer.input_dtype, DType) and isinstance(layer.meta_dtype, DType)
    assert layer.compute_dtype == DTYPE_TO_TORCH[layer.input_dtype.value]
    assert isinstance(layer.scaled_activations, bool) and isinstance(layer.data_contiguous, bool)
    for key, t in zip(("W_q", "scales", "zeros"), layer.get_tensor_args()):
        assert t.dtype == want[key].dtype and t.shape == want[key].shape
        assert torch.equal(t.contiguous().reshape(-1).view(torch.uint8), want[key].contiguous().reshape(-1).view(torch.uint8)), key
    if "bias" in want:
        assert torch.equal(layer.bias, want["bias"])
    else:
        assert layer.bias is None
    # live (helper-overridden) modes differ from the stale metadata exactly where the reference's do
    live = [int(v) for v in z["meta_args"]]
    assert [i for i in range(12) if live[i] != meta[i]] in ([], [9], [10], [9, 10]), (live, meta)


def test_state_dict_keys_and_roundtrip():
    z, info = load("a16w4_gs128_fp16_fma_bias")
    sd, N, K = reference_state_dict(z, info)
    layer = GemLiteLinear.__new__(GemLiteLinear)
    torch.nn.Module.__init__(layer)
    layer.load_state_dict(dict(sd))
    # what pack() registers (core.py:502-517): parameters, so that nn.Module.state_dict() saves them
    for k in ("W_q", "scales", "zeros", "bias"):
        setattr(layer, k, torch.nn.Parameter(getattr(layer, k), requires_grad=False))
    layer.metadata = torch.nn.Parameter(torch.tensor(layer.get_meta_args(), dtype=torch.int32), requires_grad=False)
    layer.orig_shape = torch.nn.Parameter(torch.tensor([N, K], dtype=torch.int32), requires_grad=False)
    out = layer.state_dict()
    assert set(out.keys()) == {"W_q", "bias", "scales", "zeros", "metadata", "orig_shape"}      # core.py:301-308
    buf = io.BytesIO()
    torch.save(out, buf)
    buf.seek(0)
    back = torch.load(buf)
    again = GemLiteLinear.__new__(GemLiteLinear)
    torch.nn.Module.__init__(again)
    again.load_state_dict(back)
    assert again.get_meta_args() == layer.get_meta_args() and (again.out_features, again.in_features) == (N, K)
    for a, b in zip(again.get_tensor_args(), layer.get_tensor_args()):
        assert torch.equal(a, b)
    assert torch.equal(again.bias, layer.bias)
