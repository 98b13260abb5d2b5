"""Host logic of pack() and the helper processors on the CPU: every layer fixture of the reference is rebuilt with the
PRODUCT code (gemlite_b200.core / helper: validation, mode resolution, scale / zero reshaping, the MX weight quantiser,
metadata incl. the stale-`metadata` quirk), with ONE substitution -- the CUDA bit-packer (glb_pack_cols, validated
bit-exactly on the GPU by tests/test_gpu_parity.py::test_pack_*) is replaced by the oracle's packer, because the product
has no CPU path.  Packed tensors, strides, dtypes, shapes and both metadata lists must equal the reference's bit for bit.
This is the same construction the GPU fixture test performs (`build_gpu_layer`), so a host-side regression shows up here
without a GPU."""
import numpy as np
import pytest
import torch

import gemlite_b200.bitpack as BP
import gemlite_b200.core as C
from golden_io import from_bits, layer_cases, load, tdtype
from oracle import gemlite_oracle as O


def _oracle_packer(W_q, W_nbits, packing_bitwidth=32, transpose=True):
    words, eps = O.pack_over_cols(W_q.cpu().numpy(), W_nbits, packing_bitwidth, transpose)
    return torch.from_numpy(np.ascontiguousarray(words)), eps


@pytest.fixture()
def cpu_build(monkeypatch):
    import test_gpu_parity as T
    for mod in (BP, C):
        for n in dir(mod):
            if n.startswith("pack_weights_over_cols"):
                monkeypatch.setattr(mod, n, _oracle_packer)
    monkeypatch.setattr(T, "DEV", "cpu")
    return T.build_gpu_layer


@pytest.mark.parametrize("name", layer_cases())
def test_product_host_logic_reproduces_reference_pack(name, cpu_build):
    z, info = load(name)
    layer = cpu_build(name, z, info)
    assert layer.get_meta_args() == [int(v) for v in z["meta_args"]]
    assert layer.metadata.tolist() == [int(v) for v in z["metadata_param"]]
    assert list(layer.W_q.shape) == info["W_q_shape"]
    assert list(layer.W_q.stride()) == [int(v) for v in z["packed_W_q_strides"]]
    wq = layer.W_q.data.contiguous()
    wq_bits = wq.view(torch.uint8).numpy() if wq.dtype.is_floating_point else wq.numpy()
    assert np.array_equal(wq_bits, z["packed_W_q"]), "packed W_q differs from the reference"
    sdt, zdt = tdtype(info["scales_dtype"]), tdtype(info["zeros_dtype"])
    assert layer.scales.dtype == sdt and layer.zeros.dtype == zdt
    assert list(layer.scales.shape) == info["scales_shape"] and list(layer.zeros.shape) == info["zeros_shape"]
    assert torch.equal(layer.scales.data.contiguous(), from_bits(z["packed_scales"], sdt).view(layer.scales.shape))
    assert torch.equal(layer.zeros.data.contiguous(), from_bits(z["packed_zeros"], zdt).view(layer.zeros.shape))
    assert (layer.out_features, layer.in_features) == (info["W_q_shape"][1], info["W_q_shape"][0] * layer.elements_per_sample)
