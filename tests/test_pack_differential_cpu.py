"""Differential test on the CPU: the PRODUCT's `GemLiteLinear.pack()` host logic (mode resolution, scale / zero
reshaping, packing word choice, placeholders, metadata) against the oracle's independent restatement `prepare()`
(both follow gemlite/core.py:336-519) on seeded random configurations beyond the reference fixtures.  The CUDA bit-packer
is replaced by the oracle's packer for the duration of the test (the product has no CPU path; the CUDA packer itself is
checked bit-exactly on the GPU)."""
import itertools

import numpy as np
import pytest
import torch

import gemlite_b200.bitpack as BP
import gemlite_b200.core as C
from gemlite_b200 import DType, GemLiteLinear
from gemlite_b200.dtypes import TORCH_TO_DTYPE
from oracle import gemlite_oracle as O


def _oracle_packer(W_q, W_nbits, packing_bitwidth=32, transpose=True):
    words, eps = O.pack_over_cols(W_q.cpu().numpy(), W_nbits, packing_bitwidth, transpose)
    return torch.from_numpy(np.ascontiguousarray(words)), eps


@pytest.fixture()
def cpu_packer(monkeypatch):
    for mod in (BP, C):
        for n in dir(mod):
            if n.startswith("pack_weights_over_cols"):
                monkeypatch.setattr(mod, n, _oracle_packer)


def _bits(t):
    t = t.contiguous()
    return t.reshape(-1).view(torch.uint8) if t.numel() else t.reshape(-1)


CASES = list(itertools.product((8, 4, 2, 1), (torch.float16, torch.bfloat16), ("grouped", "channelwise"),
                               ("tensor_fma", "tensor_nofma", "none", "scalar"), (None, 8, 16)))


@pytest.mark.parametrize("idx", range(0, len(CASES), 3))          # every third combination: 64 configurations
def test_pack_matches_oracle_prepare(idx, cpu_packer):
    nbits, dtype, grouping, zero_kind, pw = CASES[idx]
    g = torch.Generator().manual_seed(1000 + idx)
    N, K = 32 * (1 + idx % 3), 128 * (1 + idx % 2)
    gs = K if grouping == "channelwise" else 64
    W_q = torch.randint(0, 2 ** nbits, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    ng = N * K // gs
    scales = (torch.rand(ng, 1, generator=g) * 0.01 + 0.001).to(dtype)
    if zero_kind.startswith("tensor"):
        zeros = (torch.rand(ng, 1, generator=g) * (2 ** nbits - 1)).to(dtype)
    elif zero_kind == "scalar":
        zeros = 2 ** (nbits - 1)
    else:
        zeros = None
    fma = zero_kind != "tensor_nofma"
    bias = (torch.randn(N, generator=g) / 10).to(dtype) if idx % 2 else None
    gd = TORCH_TO_DTYPE[dtype]
    layer = GemLiteLinear(nbits, gs, K, N, gd, gd)
    layer.pack(W_q.clone(), scales.clone(), zeros.clone() if isinstance(zeros, torch.Tensor) else zeros,
               None if bias is None else bias.clone(), fma_mode=fma, packing_bitwidth=pw)
    p = O.prepare(W_q, scales, zeros, W_nbits=nbits, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], bias=bias, fma_mode=fma,
                  packing_bitwidth=32 if pw is None else pw)
    assert layer.get_meta_args() == p.meta_args(), (CASES[idx], layer.get_meta_args(), p.meta_args())
    assert layer.metadata.tolist() == p.meta_args()
    for name, a, b in (("W_q", layer.W_q.data, p.W_q), ("scales", layer.scales.data, p.scales), ("zeros", layer.zeros.data, p.zeros)):
        assert a.dtype == b.dtype and a.shape == b.shape, (CASES[idx], name, a.dtype, b.dtype, a.shape, b.shape)
        if a.dim() == 2 and a.numel():
            assert a.stride() == b.stride(), (CASES[idx], name, a.stride(), b.stride())
        assert torch.equal(_bits(a), _bits(b)), (CASES[idx], name)
    if bias is None:
        assert layer.bias is None
    else:
        assert torch.equal(layer.bias.data, p.bias)
