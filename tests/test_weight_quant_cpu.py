"""`WeightQuantizerMXFP` (load-time torch code here as in the reference, quant_utils.py:70-225) against
outputs of the reference's own quantiser (tests/golden/weight_quant_mx.npz, written by make_golden.py): MXFP8, MXFP4 and
NVFP4, with and without the window_size scale search, values and e2m1 indices -- bit for bit, and `dequantize`."""
import numpy as np
import torch

from gemlite_b200.quant_utils import WeightQuantizerMXFP
from golden_io import from_bits, load


def _bytes(t):
    if isinstance(t, np.ndarray):
        return np.ascontiguousarray(t).reshape(-1).view(np.uint8)
    return t.contiguous().reshape(-1).view(torch.uint8).numpy()


def test_weight_quantiser_matches_reference_bit_for_bit():
    z, info = load("weight_quant_mx")
    q = WeightQuantizerMXFP(compute_dtype=torch.bfloat16, device="cpu")
    seen = 0
    for tag, dt in (("bf16", torch.bfloat16), ("f16", torch.float16)):
        W = from_bits(z[f"W_{tag}"], dt)
        for key in info["kernels_ok"]:
            if f"_{tag}_" not in key:
                continue
            parts = key.split("_")
            fn, index = parts[0] + "_" + parts[1], key.endswith("i1")
            if fn == "quantize_mxfp8":
                Wq, sc = q.quantize_mxfp8(W.clone(), index=index)
            else:
                ws = int([p for p in parts if p[0] == "w" and p[1:].isdigit()][0][1:])
                Wq, sc = getattr(q, fn)(W.clone(), window_size=ws, index=index)
            assert np.array_equal(_bytes(Wq), _bytes(z["q_" + key])), key
            assert np.array_equal(_bytes(sc), _bytes(z["s_" + key])), key
            seen += 1
        Wq, sc = q.quantize_mxfp4(W.clone(), window_size=1, index=True)
        deq = q.dequantize(Wq, sc, shape=W.shape, dtype=dt)
        assert torch.equal(deq, from_bits(z[f"deq_mxfp4_{tag}"], dt).view(W.shape))
    assert seen == len(info["kernels_ok"]) == 28
