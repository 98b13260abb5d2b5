"""GPU parity: the CUDA path (through the public API -> C ABI) against (a) the golden fixtures written
by the reference itself and (b) the CPU oracle on seeded inputs.  Integer / byte work is bit-exact;
floating point uses the tolerances stated next to each check."""
import os

import numpy as np
import pytest
import torch

import gemlite_b200
from gemlite_b200 import DType, GemLiteLinear, _cabi
from gemlite_b200 import bitpack as B
from gemlite_b200.helper import (A16W8_INT8, A16W4_HQQ_INT, A16W8_FP8, A8W8_INT8_dynamic, A8W8_FP8_dynamic,
                                 A16W4_MXFP, A16W8_MXFP, A16W158_INT, A8W8_MXFP_dynamic, A8W4_MXFP_dynamic,
                                 A4W4_MXFP_dynamic, A4W4_NVFP_dynamic)
from gemlite_b200.quant_utils import scale_activations_per_token
from oracle import gemlite_oracle as O
from oracle import recipes as R
from golden_io import MX_DYNAMIC, from_bits, layer_cases, load, tdtype

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MXV_NAME = "gemv_mxfp4"


def a8_kernel_name(kind, M, K=1024, N=1024):
    if M <= 16:                 # streaming kernel (gl_gemv_a8.cu), validated in round 2
        return f"gemv_a8w8_{kind}"
    if M <= 128 and K in (1024, 2048, 4096, 8192) and N % 128 == 0:     # batched decode (gl_gemm_a8_mid.cu)
        return f"gemm_a8w8_{kind}_mma_mid"
    return f"gemm_a8w8_{kind}_tcgen05_2cta" if M > 128 else f"gemm_a8w8_{kind}_tcgen05"

# |y_cuda - y_oracle| <= REL * max|y_oracle|: one output ulp of headroom plus fp32 summation-order noise
REL = {torch.float16: 2.0e-3, torch.bfloat16: 1.6e-2, torch.float32: 2e-5}
# the decode kernel factors scale/zero per group (y += s*sum(x q) + z'*sum(x)) instead of rounding every
# dequantised weight to fp16 first; that differs from the oracle by weight-rounding noise (<= 2^-11 rel.
# per weight, averaging out over K) -- bound stated against max|y|:
REL_DECODE = {torch.float16: 3.0e-3, torch.bfloat16: 2.0e-2}


def relerr(y, yref):
    y, yref = y.double().cpu(), yref.double().cpu()
    return ((y - yref).abs().max() / (yref.abs().max() + 1e-30)).item()


def oracle_from_layer(layer):
    """Mirror a packed GPU layer into the oracle (same tensors, same metadata) -- used when the quantisation
    recipe itself ran on the GPU: torch's CUDA `x / 127` is a reciprocal multiply and can differ from the CPU
    divide by one ulp, which is a property of torch, not of the path under test."""
    p = O.Packed()
    p.W_q = layer.W_q.data.cpu()
    p.scales, p.zeros = layer.scales.data.cpu(), layer.zeros.data.cpu()
    p.bias = None if layer.bias is None else layer.bias.data.cpu()
    (sa, p.W_nbits, p.group_size, p.unpack_mask, p.elements_per_sample, p.input_dtype, p.output_dtype, p.acc_dtype,
     p.meta_dtype, p.channel_scale_mode, p.W_group_mode, dc) = layer.get_meta_args()
    p.scaled_activations, p.data_contiguous = bool(sa), bool(dc)
    p.in_features, p.out_features = layer.in_features, layer.out_features
    return p


def gen_int(N, K, nbits, gs, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    W_q = torch.randint(0, 2 ** nbits, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    ng = N * K // gs
    scales = (torch.rand(ng, 1, generator=g) * 0.01 + 0.001).to(dtype)
    zeros = (torch.rand(ng, 1, generator=g) * (2 ** nbits - 1)).to(dtype)
    return W_q, scales, zeros


def xgen(M, K, dtype, seed):
    return (torch.randn(M, K, generator=torch.Generator().manual_seed(seed)) / 10.0).to(dtype)


# ------------------------------------------------------------------------------------------------
def test_native_library_is_loaded():
    lib = _cabi.load()
    assert lib.glb_version() == 1
    assert "libgemlite_b200.so" in open("/proc/self/maps").read()


@pytest.mark.parametrize("nbits", [8, 4, 2, 1])
@pytest.mark.parametrize("pw", [8, 16, 32, 64])
def test_pack_unpack_bit_exact(nbits, pw):
    if pw < nbits:
        pytest.skip("word narrower than a code")
    g = torch.Generator().manual_seed(nbits * 100 + pw)
    for (rows, cols) in [(16, 128), (96, 4096), (33, 256)]:
        W_q = torch.randint(0, 2 ** nbits, (rows, cols), generator=g, dtype=torch.int32).to(torch.uint8)
        for tr in (True, False):
            ref, eps = O.pack_over_cols(W_q.numpy(), nbits, pw, tr)
            out, eps2 = B.pack_weights_over_cols(W_q.to(DEV), nbits, pw, tr)
            assert eps == eps2 and tuple(out.shape) == ref.shape
            assert np.array_equal(out.cpu().numpy(), np.ascontiguousarray(ref)), (nbits, pw, tr, rows, cols)
            back = B.unpack_over_cols(out, nbits, cols, transpose=tr)
            assert torch.equal(back.cpu(), W_q)


def test_pack_matches_reference_fixture():
    z, _ = load("bitpack")
    for nbits in (8, 4, 2, 1):
        W_q = torch.from_numpy(z[f"in_b{nbits}"]).to(DEV)
        for pw in (8, 16, 32, 64):
            if pw < nbits:
                continue
            for tr in (1, 0):
                out, _ = B.pack_weights_over_cols(W_q, nbits, pw, bool(tr))
                key = f"triton_b{nbits}_w{pw}_t{tr}"
                ref = z[key] if key in z.files else z[f"torch_b{nbits}_w{pw}_t{tr}"]
                assert np.array_equal(out.cpu().numpy(), ref), key


# ------------------------------------------------------------------------------------------------
def build_gpu_layer(name, z, info):
    """Same construction as tests/golden/make_golden.py, on the GPU with gemlite_b200."""
    if name in MX_DYNAMIC:
        # the reference's own quantised weights and scales (pre-quantised branch of from_weights, helper.py:668-705)
        kind, post = MX_DYNAMIC[name]
        N, K = from_bits(z["in_W"], torch.bfloat16).shape
        gs = 16 if kind == "a4w4_nv" else 32
        if kind == "a8w8":
            W_q = from_bits(z["q_codes"], torch.float8_e4m3fn).view(N, K)
        else:
            W_q = torch.from_numpy(z["q_codes"]).view(N, K)
        sc = from_bits(z["q_scales"], torch.float8_e4m3fn if kind == "a4w4_nv" else torch.float8_e8m0fnu).view(N, K // gs)
        proc = {"a8w8": lambda: A8W8_MXFP_dynamic(device=DEV, dtype=torch.bfloat16, post_scale=post),
                "a8w4": lambda: A8W4_MXFP_dynamic(device=DEV, dtype=torch.bfloat16, post_scale=post),
                "a4w4_mx": lambda: A4W4_MXFP_dynamic(device=DEV, dtype=torch.bfloat16),
                "a4w4_nv": lambda: A4W4_NVFP_dynamic(device=DEV, dtype=torch.bfloat16)}[kind]()
        return proc.from_weights(W_q, scales=sc)
    if name.startswith("a16w") and "in_W_q" in z.files:
        parts = name.split("_")
        nbits = int(parts[0][4:])
        dtype = torch.bfloat16 if "bf16" in name else torch.float16
        gs = 256 if "channelwise" in name else int([p for p in parts if p.startswith("gs")][0][2:])
        W_q = torch.from_numpy(z["in_W_q"]).to(DEV)
        scales = from_bits(z["in_scales"], dtype).to(DEV)
        zeros = from_bits(z["in_zeros"], dtype).to(DEV)
        bias = from_bits(z["in_bias"], dtype).to(DEV) if "in_bias" in z.files else None
        N, K = W_q.shape
        gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
        layer = GemLiteLinear(nbits, gs, K, N, gd, gd)
        layer.pack(W_q, scales, zeros, bias, fma_mode="nofma" not in name,
                   packing_bitwidth=8 if name.endswith("pw8") else None)
        return layer
    if name.startswith("a16w8_int8") or name.startswith("a8w8_"):
        # feed the reference's own quantised weights / scales (pre-quantised branch of from_weights): the float
        # quantisation recipe is torch code whose CUDA and CPU divides differ by an ulp (see oracle_from_layer)
        wdt = tdtype(info["W_q_dtype"])
        K, N = info["W_q_shape"]
        W_q = from_bits(z["packed_W_q"], wdt).view(K, N).t().contiguous()
        sc = from_bits(z["packed_scales"], torch.float32).view(N, 1)
        if name.startswith("a16w8_int8"):
            return A16W8_INT8(device=DEV, dtype=torch.float16, post_scale=name.endswith("postscale")).from_weights(W_q, scales=sc)
        proc = A8W8_FP8_dynamic(device=DEV, dtype=torch.float16) if "fp8" in name else \
            A8W8_INT8_dynamic(device=DEV, dtype=torch.float16)
        return proc.from_weights(W_q, scales=sc)
    if name == "int8_w4_scalar_zero":
        W_q = torch.from_numpy(z["in_W_q"]).to(DEV)
        sc = from_bits(z["in_scales"], torch.float16).to(DEV)
        N, K = W_q.shape
        layer = GemLiteLinear(4, K, K, N, DType.INT8, DType.FP32, scaled_activations=False)
        layer.pack(W_q, scales=sc, zeros=7, bias=None)
        return layer
    if "mxfp" in name:
        W = from_bits(z["in_W"], torch.bfloat16)
        lin = torch.nn.Linear(W.shape[1], W.shape[0], bias=False, dtype=torch.bfloat16, device=DEV)
        lin.weight.data = W.to(DEV)
        proc = (A16W4_MXFP if name.startswith("a16w4") else A16W8_MXFP)(device=DEV, dtype=torch.bfloat16)
        return proc.from_linear(lin, del_orig=False)
    raise KeyError(name)


@pytest.mark.parametrize("name", layer_cases())
def test_against_reference_fixtures(name):
    """pack() reproduces the reference's tensors/metadata bit for bit; forward() matches the reference
    kernels' recorded outputs (where the CPU interpreter executes them faithfully) and the oracle."""
    z, info = load(name)
    layer = build_gpu_layer(name, z, info)
    assert layer.get_meta_args() == [int(v) for v in z["meta_args"]]
    assert list(layer.W_q.shape) == info["W_q_shape"]
    assert list(layer.W_q.stride()) == [int(v) for v in z["packed_W_q_strides"]]
    wq = layer.W_q.data.contiguous().cpu()
    wq_bits = wq.view(torch.uint8).numpy() if wq.dtype.is_floating_point else wq.numpy()
    assert np.array_equal(wq_bits, z["packed_W_q"]), "packed W_q differs from the reference"
    sdt, zdt = tdtype(info["scales_dtype"]), tdtype(info["zeros_dtype"])
    assert layer.scales.dtype == sdt and layer.zeros.dtype == zdt
    assert list(layer.scales.shape) == info["scales_shape"] and list(layer.zeros.shape) == info["zeros_shape"]
    assert torch.equal(layer.scales.data.contiguous().cpu(), from_bits(z["packed_scales"], sdt).view(layer.scales.shape))
    assert torch.equal(layer.zeros.data.contiguous().cpu(), from_bits(z["packed_zeros"], zdt).view(layer.zeros.shape))
    assert layer.metadata.tolist() == [int(v) for v in z["metadata_param"]]

    # oracle layer from the same inputs
    from test_oracle_golden import build_oracle_layer, _interpreter_untrustworthy
    p = build_oracle_layer(name, z, info)
    xdt = tdtype(info["x_dtype"])
    out_t = O.TORCH_OF[p.output_dtype]
    i = 0
    while f"x{i}" in z.files:
        x = from_bits(z[f"x{i}"], xdt)
        M = x.shape[0]
        y_or = O.forward(p, x)
        for mt in (["GEMV_REVSPLITK", "GEMV", "GEMV_SPLITK"] if M == 1 else []) + ["GEMM_SPLITK", "GEMM", None]:
            y = layer.forward_manual(x.to(DEV), mt) if mt else layer(x.to(DEV))
            assert y.dtype == out_t and tuple(y.shape) == (M, p.out_features)
            kern = _cabi.last_kernel()
            tol = REL_DECODE[out_t] if kern == "gemv_w4_mma" else REL[out_t]
            if p.scaled_activations and xdt != torch.int8:
                tol = max(tol, 2e-2)   # an activation code may round the other way (div.full.f32 vs IEEE divide)
            e = relerr(y, y_or)
            assert e <= tol, (name, mt, kern, e, tol)
            # reference's own recorded output for the fp32-accumulating kernels
            key = f"y_GEMM_SPLITK_M{M}"
            if key in z.files and not _interpreter_untrustworthy(name, p):
                e2 = relerr(y, from_bits(z[key], out_t))
                assert e2 <= max(tol, 2.5e-3), (name, mt, "vs reference GEMM_SPLITK", e2)
        i += 1


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", ["fma", "nofma", "sym"])
@pytest.mark.parametrize("gs", [32, 64, 128])
def test_a16w4_decode_and_small_batch_vs_oracle(dtype, mode, gs):
    N, K = 256, 1024
    W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=gs + (7 if mode == "fma" else 11))
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(3)) / 10).to(dtype)
    layer = GemLiteLinear(4, gs, K, N, gd, gd)
    z_in = None if mode == "sym" else zeros
    layer.pack(W_q.to(DEV), scales.to(DEV), None if z_in is None else z_in.to(DEV), bias.to(DEV), fma_mode=(mode == "fma"))
    p = O.prepare(W_q, scales, z_in, W_nbits=4, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], bias=bias, fma_mode=(mode == "fma"))
    assert layer.W_group_mode == p.W_group_mode == {"fma": 4, "nofma": 3, "sym": 2}[mode]
    for M in (1, 2, 5, 8, 16, 33):
        x = xgen(M, K, dtype, 50 + M)
        y = layer(x.to(DEV))
        kern = _cabi.last_kernel()
        tol = REL_DECODE[dtype] if kern == "gemv_w4_mma" else REL[dtype]
        e = relerr(y, O.forward(p, x))
        assert e <= tol, (dtype, mode, gs, M, kern, e)
    assert layer(xgen(1, K, dtype, 1).to(DEV)).shape == (1, N)
    # leading batch dims are preserved (core.py:151-152)
    assert layer(xgen(6, K, dtype, 2).view(2, 3, K).to(DEV)).shape == (2, 3, N)


# ------------------------------------------------------------------------------------------------
# Prefill: the CTA-pair tcgen05 kernel (gl_gemm_tc2.cu: 2-CTA MMA, dequantised weights in tensor memory)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", ["fma", "nofma", "sym"])
@pytest.mark.parametrize("gs", [32, 64, 128])
def test_a16w4_prefill_pair_kernel_vs_oracle(dtype, mode, gs):
    N, K = 512, 1024
    W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=gs + (17 if mode == "fma" else 23))
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(4)) / 10).to(dtype)
    layer = GemLiteLinear(4, gs, K, N, gd, gd)
    z_in = None if mode == "sym" else zeros
    layer.pack(W_q.to(DEV), scales.to(DEV), None if z_in is None else z_in.to(DEV), bias.to(DEV), fma_mode=(mode == "fma"))
    p = O.prepare(W_q, scales, z_in, W_nbits=4, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], bias=bias, fma_mode=(mode == "fma"))
    for M in (129, 256, 300, 700):          # ragged last token block, one and several token blocks
        x = xgen(M, K, dtype, 90 + M)
        y = layer(x.to(DEV))
        assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta", _cabi.last_kernel()   # group_size % 32 == 0, K % 128 == 0
        e = relerr(y, O.forward(p, x))
        assert e <= REL[dtype], (dtype, mode, gs, M, e)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_prefill_pair_kernel_dequant_is_bit_exact(dtype):
    """x = identity => y[m, n] is the dequantised weight W[n, m] itself (one non-zero product per output, fp32
    accumulate): the tensor-memory operand must hold exactly the reference's dequantize() values (utils.py:57-89)."""
    N, K, gs = 512, 512, 64
    W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=5)
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    layer = GemLiteLinear(4, gs, K, N, gd, gd).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV))
    p = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype])
    x = torch.eye(K, dtype=dtype)
    y = layer(x.to(DEV))
    assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta"
    assert torch.equal(y.cpu(), O.forward(p, x))


def test_prefill_pair_kernel_many_tiles_per_pair():
    """More tiles than CTA pairs (100 > 74): exercises the accumulator hand-over between consecutive tiles and the
    ring phases; checked on the device against torch fp32 dequant + matmul."""
    M, N, K, gs = 2560, 2560, 512, 128
    W_q, scales, zeros = gen_int(N, K, 4, gs, torch.float16, seed=9)
    layer = A16W4_HQQ_INT(device=DEV, dtype=torch.float16).from_weights(W_q, scales, zeros)
    x = xgen(M, K, torch.float16, 77).to(DEV)
    y = layer(x)
    assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta"
    Wd = ((W_q.to(DEV).view(-1, gs).float() - zeros.to(DEV).float()) * scales.to(DEV).float()).view(N, K)
    assert relerr(y, x.float() @ Wd.t()) <= REL[torch.float16]
    assert torch.equal(layer(x), y)          # deterministic


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_pair_kernel_split_k_small_batches(dtype):
    """16 < M <= 256 on a narrow layer: the pair kernel cuts K into splits so that all SM pairs work; partial sums are
    parked in the workspace and reduced in split order by the last warp to arrive -> deterministic, counters self-clean."""
    N, K, gs = 512, 2048, 128
    W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=21)
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(5)) / 10).to(dtype)
    layer = GemLiteLinear(4, gs, K, N, gd, gd).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV), bias.to(DEV))
    p = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], bias=bias)
    for M in (17, 64, 65, 128, 200, 256):
        x = xgen(M, K, dtype, 500 + M)
        xd = x.to(DEV)
        y = layer.forward_manual(xd, "GEMM")     # by name: the automatic choice for M <= 128 is the batched-decode kernel
        assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta", (M, _cabi.last_kernel())
        assert relerr(y, O.forward(p, x)) <= REL[dtype], (dtype, M)
        for _ in range(3):                       # same workspace, counters must have returned to zero; bitwise repeatable
            assert torch.equal(layer.forward_manual(xd, "GEMM"), y)


def test_split_k_layers_of_different_width_share_one_workspace():
    """One per-stream workspace serves every layer.  Layers with different tile counts (q/o vs gate/up projections) take
    turns in the split-K range: the arrival counters live in a fixed region of the workspace, so the partial tiles one
    layer parks there can never be read as counters by the next (they could when the layout depended on the tile count)."""
    dtype, gs, K = torch.float16, 128, 1024
    layers = []
    for i, N in enumerate((4096, 2048, 256)):          # 16, 8 and 1 feature tiles: 512 / 256 / 32 bytes of counters
        W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=40 + i)
        layer = GemLiteLinear(4, gs, K, N, DType.FP16, DType.FP16).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV))
        p = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=gs, in_features=K, out_features=N)
        layers.append((layer, p))
    for M in (32, 100):
        x = xgen(M, K, dtype, 800 + M)
        want = [O.forward(p, x) for _, p in layers]
        xd = x.to(DEV)
        first = None
        for rep in range(3):
            got = [layer.forward_manual(xd, "GEMM") for layer, _ in layers]
            assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta"
            for (layer, p), y, w in zip(layers, got, want):
                assert relerr(y, w) <= REL[dtype], (M, rep, p.out_features, relerr(y, w))
            if first is None:
                first = got
            else:
                assert all(torch.equal(a, b) for a, b in zip(first, got)), (M, rep)


def test_decode_kernel_is_used_for_the_headline_shape():
    N = K = 4096
    W_q, scales, zeros = gen_int(N, K, 4, 128, torch.float16, seed=0)
    layer = A16W4_HQQ_INT(device=DEV, dtype=torch.float16).from_weights(W_q, scales, zeros)
    assert (layer.W_group_mode, layer.channel_scale_mode, layer.elements_per_sample) == (4, 0, 8)
    assert tuple(layer.W_q.shape) == (512, 4096) and layer.W_q.dtype == torch.int32 and layer.W_q.is_contiguous()
    x = xgen(1, K, torch.float16, 5).to(DEV)
    before = _cabi.launch_count()
    y = layer(x)
    assert _cabi.last_kernel() == "gemv_w4_mma" and _cabi.launch_count() == before + 1
    # full-size check against a torch fp32 dequant + matmul on the device (tolerance: REL_DECODE)
    Wd = ((W_q.to(DEV).view(-1, 128).float() - zeros.to(DEV).float()) * scales.to(DEV).float()).view(N, K)
    yref = x.float() @ Wd.t()
    assert relerr(y, yref) <= REL_DECODE[torch.float16]
    # linearity in x (size-independent property): y(2x) == 2 y(x) up to one output ulp (power-of-two scaling
    # commutes with every rounding on the path except for x/16 terms that are subnormal in fp16)
    assert relerr(layer(2 * x), 2 * y) <= 1.5e-3
    # determinism: no atomics anywhere on the path
    assert torch.equal(layer(x), y)


def test_per_token_activation_quant_vs_oracle():
    for dtype in (torch.float16, torch.bfloat16, torch.float32):
        for wdt in (torch.int8, torch.float8_e4m3fn):
            x = xgen(37, 4096, dtype, 99)
            q, s = scale_activations_per_token(x.to(DEV), wdt)
            qo, so = O.quantize_activations_per_token(x, wdt)
            assert s.shape == (37, 1) and q.shape == x.shape and q.dtype == wdt
            # scale: div.full.f32 (<= 2 ulp) vs IEEE divide
            assert torch.allclose(s.cpu(), so, rtol=3e-7, atol=0)
            if wdt == torch.int8:
                d = (q.cpu().int() - qo.int()).abs()
                assert d.max().item() <= 1 and (d > 0).float().mean().item() < 2e-3
            else:
                mism = (q.cpu().view(torch.uint8) != qo.view(torch.uint8)).float().mean().item()
                assert mism < 2e-3
    # empty and single-row inputs
    q, s = scale_activations_per_token(torch.zeros(1, 64, dtype=torch.float16, device=DEV), torch.int8)
    assert s.item() == pytest.approx(1e-6) and int(q.abs().max()) == 0


def test_int8_matmul_bit_exact():
    """INT8 x INT8 -> INT32 is exact; with the activation codes/scales the GPU produced, the fp16 output
    must equal the oracle's bit for bit at every batch size (epilogue cvt.rn.f32.s32 * (sa*sb) -> fp16)."""
    N = K = 1024
    W = (torch.randn(N, K, generator=torch.Generator().manual_seed(8)) / 10).half()
    layer = A8W8_INT8_dynamic(device=DEV, dtype=torch.float16).from_weights(W.clone())
    p = oracle_from_layer(layer)
    p_cpu = R.a8w8_dynamic(W, torch.float16)     # the recipe itself: at most one code / one ulp apart (CUDA vs CPU divide)
    assert (layer.W_q.data.cpu().int() - p_cpu.W_q.int()).abs().max().item() <= 1
    assert torch.allclose(layer.scales.data.cpu(), p_cpu.scales, rtol=3e-7, atol=0)
    assert (p.W_group_mode, p.channel_scale_mode, p.meta_dtype) == (p_cpu.W_group_mode, p_cpu.channel_scale_mode, p_cpu.meta_dtype)
    for M in (1, 3, 16, 64, 65, 200, 300, 777):
        x = xgen(M, K, torch.float16, M)
        xq, sx = scale_activations_per_token(x.to(DEV), torch.int8)
        y = layer(x.to(DEV))
        # M > 128 runs on the CTA-pair kernel (gl_gemm_a8_tc2.cu), below that on the single-CTA one: same bits required
        assert _cabi.last_kernel() == a8_kernel_name("int8", M, K, N), (M, _cabi.last_kernel())
        y_or = O.forward(p, xq.cpu(), scales_x=sx.cpu())
        assert torch.equal(y.cpu(), y_or), (M, relerr(y, y_or))
        # raw int32 accumulator through an unscaled layer
    acc_layer = GemLiteLinear(8, None, K, N, DType.INT8, DType.INT32)
    Wi = torch.randint(-128, 128, (N, K), generator=torch.Generator().manual_seed(1), dtype=torch.int32).to(torch.int8)
    acc_layer.pack(Wi.to(DEV), None, None)
    for Mi in (9, 300):          # single-CTA and CTA-pair kernel
        xi = torch.randint(-128, 128, (Mi, K), generator=torch.Generator().manual_seed(2), dtype=torch.int32).to(torch.int8)
        got = acc_layer(xi.to(DEV))
        assert got.dtype == torch.int32 and torch.equal(got.cpu(), xi.int() @ Wi.int().t())


def test_fp8_dynamic_and_weight_only_paths():
    N = K = 512
    W = (torch.randn(N, K, generator=torch.Generator().manual_seed(18)) / 10).half()
    layer = A8W8_FP8_dynamic(device=DEV, dtype=torch.float16).from_weights(W.clone())
    p = oracle_from_layer(layer)
    for M in (1, 4, 70, 300, 520):
        x = xgen(M, K, torch.float16, 40 + M)
        xq, sx = scale_activations_per_token(x.to(DEV), torch.float8_e4m3fn)
        y = layer(x.to(DEV))
        assert _cabi.last_kernel() == a8_kernel_name("fp8", M, K, N), (M, _cabi.last_kernel())
        assert relerr(y, O.forward(p, xq.cpu(), scales_x=sx.cpu())) <= REL[torch.float16]
    wl = A16W8_FP8(device=DEV, dtype=torch.float16).from_weights(W.clone())
    pw = oracle_from_layer(wl)
    assert (pw.W_group_mode, pw.channel_scale_mode) == (2, 0)
    x = xgen(3, K, torch.float16, 77)
    assert relerr(wl(x.to(DEV)), O.forward(pw, x)) <= REL[torch.float16]


def test_a8w8_decode_kernel_vs_oracle():
    """gl_gemv_a8.cu: M <= 16, every scaling mode / output type of the A8W8 path; INT8 bit-exact, FP8 within REL."""
    N, K = 1024, 2048
    g = torch.Generator().manual_seed(31)
    W = (torch.randn(N, K, generator=g) / 10).half()
    bias = (torch.randn(N, generator=g) / 10)
    for odt in (torch.float16, torch.bfloat16):
        for proc, wdt, kind in ((A8W8_INT8_dynamic, torch.int8, "int8"), (A8W8_FP8_dynamic, torch.float8_e4m3fn, "fp8")):
            for b in (None, bias.to(odt)):
                layer = proc(device=DEV, dtype=odt).from_weights(W.clone().to(odt), bias=None if b is None else b.clone())
                p = oracle_from_layer(layer)
                for M in (1, 2, 5, 8, 9, 16):
                    x = xgen(M, K, odt, 300 + M)
                    xq, sx = scale_activations_per_token(x.to(DEV), wdt)
                    y = layer(x.to(DEV))
                    assert _cabi.last_kernel() == f"gemv_a8w8_{kind}", (M, _cabi.last_kernel())
                    y_or = O.forward(p, xq.cpu(), scales_x=sx.cpu())
                    if kind == "int8":
                        assert torch.equal(y.cpu(), y_or), (odt, M, b is not None, relerr(y, y_or))
                    else:
                        assert relerr(y, y_or) <= REL[odt], (odt, M, b is not None, relerr(y, y_or))
    # raw INT32 accumulators, every M of both column tiles, ragged N / K multiples
    for (Ni, Ki) in ((48, 512), (1024, 4096)):
        acc_layer = GemLiteLinear(8, None, Ki, Ni, DType.INT8, DType.INT32)
        Wi = torch.randint(-128, 128, (Ni, Ki), generator=torch.Generator().manual_seed(1), dtype=torch.int32).to(torch.int8)
        acc_layer.pack(Wi.to(DEV), None, None)
        for Mi in range(1, 17):
            xi = torch.randint(-128, 128, (Mi, Ki), generator=torch.Generator().manual_seed(2 + Mi), dtype=torch.int32).to(torch.int8)
            got = acc_layer(xi.to(DEV))
            if Ni % 128 == 0:
                assert _cabi.last_kernel() == "gemv_a8w8_int8", (Ni, Ki, Mi, _cabi.last_kernel())
            assert got.dtype == torch.int32 and torch.equal(got.cpu(), xi.int() @ Wi.int().t()), (Ni, Ki, Mi)


def test_mx_activation_quantisers_bit_exact():
    """glb_quant_act_mx against (a) the reference's own outputs on the edge-case fixture and (b) the oracle on seeded
    inputs: codes AND scales bit for bit (fp32 arithmetic with IEEE division on both sides)."""
    from gemlite_b200.quant_utils import scale_activations_mxfp8, scale_activations_mxfp4, scale_activations_nvfp4
    gpu = {"mxfp8": scale_activations_mxfp8, "mxfp4": scale_activations_mxfp4, "nvfp4": scale_activations_nvfp4}
    cpu = {"mxfp8": O.quantize_activations_mxfp8, "mxfp4": O.quantize_activations_mxfp4, "nvfp4": O.quantize_activations_nvfp4}
    z, info = load("actq_mx")
    for tag, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        for M in (1, 5, 33):
            x = from_bits(z[f"x_{tag}_M{M}"], dt)
            for kind in gpu:
                q, sc = gpu[kind](x.to(DEV))
                assert _cabi.last_kernel() == "act_quant_mx"
                key = f"{kind}_{tag}_M{M}"
                assert np.array_equal(q.cpu().view(torch.uint8).numpy(), z["q_" + key]), key
                assert np.array_equal(sc.cpu().view(torch.uint8).numpy(), z["s_" + key]), key
    for dt in (torch.float16, torch.bfloat16, torch.float32):
        for (M, K) in ((1, 4096), (7, 1024), (64, 2048), (130, 512)):
            x = xgen(M, K, dt, 900 + M) * (1 + 50 * torch.rand(M, 1, generator=torch.Generator().manual_seed(M)))
            for kind in gpu:
                q, sc = gpu[kind](x.to(DEV))
                qo, so = cpu[kind](x)
                assert q.shape == qo.shape and sc.shape == so.shape and q.dtype == qo.dtype and sc.dtype == so.dtype
                assert torch.equal(q.cpu().view(torch.uint8), qo.view(torch.uint8)), (kind, dt, M, K)
                assert torch.equal(sc.cpu().view(torch.uint8), so.view(torch.uint8)), (kind, dt, M, K)
    # 3-d input: flattened like the reference's *_triton_v2 functions
    q, sc = scale_activations_mxfp4(xgen(6, 256, torch.float16, 5).view(2, 3, 256).to(DEV))
    assert tuple(q.shape) == (6, 128) and tuple(sc.shape) == (32, 8)


def test_mx_dynamic_layers_vs_oracle():
    """A8W8 / A8W4 / A4W4 MXFP and A4W4 NVFP with activations quantised on the fly, larger shapes than the fixtures,
    through from_linear (GPU weight quantiser) -> forward, against the oracle fed with the layer's own tensors."""
    N, K = 512, 1024
    W = (torch.randn(N, K, generator=torch.Generator().manual_seed(61)) / 10).to(torch.bfloat16)
    cases = [(A8W8_MXFP_dynamic, {"post_scale": True}), (A8W8_MXFP_dynamic, {"post_scale": False}),
             (A8W4_MXFP_dynamic, {"post_scale": True}), (A8W4_MXFP_dynamic, {"post_scale": False}),
             (A4W4_MXFP_dynamic, {}), (A4W4_NVFP_dynamic, {})]
    for cls, kw in cases:
        for dt in (torch.bfloat16, torch.float16):
            lin = torch.nn.Linear(K, N, bias=True, dtype=dt, device=DEV)
            lin.weight.data = W.to(device=DEV, dtype=dt)
            layer = cls(device=DEV, dtype=dt, **kw).from_linear(lin, del_orig=False)
            p = oracle_from_layer(layer)
            for M in (1, 5, 40, 130):
                x = xgen(M, K, dt, 700 + M)
                y = layer(x.to(DEV))
                # MXFP4 weights + MX micro-scaled activations in bf16: exact bf16 expansion of the quantised activations, then
                # the MXFP4 tensor-core / streaming kernels (gl_api.cu: mx_expand_view); everything else: general kernel
                expanded = (dt == torch.bfloat16 and layer.channel_scale_mode == 4 and layer.W_nbits == 4
                            and cls is not A4W4_NVFP_dynamic)
                want = ("gemv_mxfp4" if M <= 4 else "gemm_a16w4_mxfp4_tcgen05_2cta") if expanded else "generic_cuda_core"
                # both operands of one element type: the tensor core applies the group scales itself (gl_gemm_bs.cu)
                native = {A8W8_MXFP_dynamic: "gemm_mxfp8_block_scaled_tcgen05_2cta", A4W4_MXFP_dynamic: "gemm_mxfp4_block_scaled_tcgen05_2cta",
                          A4W4_NVFP_dynamic: "gemm_nvfp4_block_scaled_tcgen05_2cta"}.get(cls)
                if native and not (expanded and M <= 4):
                    want = native
                assert _cabi.last_kernel() == want, (cls.__name__, kw, dt, M, _cabi.last_kernel())
                if layer.channel_scale_mode == 2:
                    # per-token fp8 activations: the GPU quantiser divides with div.full.f32 like the reference's Triton
                    # build (codes / scales bit-equal to it: test_gpu_vs_reference.py), the CPU oracle with the IEEE
                    # divide -- a code on an exact tie can differ.  Feed the oracle the product's own codes.
                    q, sx = scale_activations_per_token(x.to(DEV), w_dtype=torch.float8_e4m3fn)
                    yo = O.forward(p, q.cpu(), scales_x=sx.cpu())
                else:
                    yo = O.forward(p, x)
                e = relerr(y, yo)
                assert e <= max(REL[dt], 2e-3), (cls.__name__, kw, dt, M, e)


def test_mxfp_weight_only_vs_oracle_and_unquantised():
    torch.manual_seed(0)
    N, K = 512, 1024
    for nbits, tol_q in ((4, 7e-4), (8, 2e-4)):   # the reference's own bars vs the unquantised layer (test_mxfp.py:50,57)
        # same data recipe as the reference test (test_mxfp.py:22-33): default Linear init / 10, x = randn / 10
        lin = torch.nn.Linear(K, N, bias=False, dtype=torch.bfloat16, device=DEV)
        lin.weight.data /= 10.0
        lin.weight.requires_grad = False
        W = lin.weight.data.cpu()
        proc = (A16W4_MXFP if nbits == 4 else A16W8_MXFP)(device=DEV, dtype=torch.bfloat16)
        layer = proc.from_linear(lin, del_orig=False)
        assert layer.W_q.numel() * layer.W_q.itemsize == N * K * nbits // 8 and not layer.scaled_activations
        p, codes, sc = R.a16wn_mxfp(W, nbits, torch.bfloat16)
        assert torch.equal(layer.scales.data.cpu(), p.scales) and list(layer.W_q.stride()) == list(p.W_q.stride())
        assert torch.equal(layer.W_q.data.cpu().contiguous().view(torch.uint8), p.W_q.contiguous().view(torch.uint8))
        for M in (1, 4, 16):
            x = (torch.randn(M, K) / 10).to(torch.bfloat16)
            y = layer(x.to(DEV))
            assert relerr(y, O.forward(p, x)) <= REL[torch.bfloat16]
            assert (lin(x.to(DEV)) - y).abs().mean().item() < tol_q


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_mxfp4_pair_kernel_vs_oracle(dtype):
    """A16W4_MXFP on the CTA-pair tcgen05 kernel (e2m1 codes -> cvt.rn.f16x2.e2m1x2 -> tensor memory): random codes
    (all 16 patterns incl. -0) and a wide band of e8m0 exponents; e2m1 x 2^e is exact in bf16 (and in fp16 inside
    its normal range), so the only difference to the oracle is fp32 accumulation order."""
    N, K = 512, 1024
    g = torch.Generator().manual_seed(11)
    codes = torch.randint(0, 16, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    lo, hi = (105, 131) if dtype == torch.bfloat16 else (117, 126)
    sc = torch.randint(lo, hi, (N, K // 32), generator=g, dtype=torch.int32).to(torch.uint8).view(torch.float8_e8m0fnu)
    bias = (torch.randn(N, generator=g) / 10).to(dtype)
    layer = A16W4_MXFP(device=DEV, dtype=dtype).from_weights(codes.to(DEV), sc.to(DEV), bias=bias.to(DEV))
    gd = O.MXFP16 if dtype == torch.float16 else O.MXBF16
    p = O.prepare(codes, sc, None, W_nbits=4, group_size=32, in_features=K, out_features=N, input_dtype=gd, output_dtype=gd,
                  bias=bias)
    assert torch.equal(layer.W_q.data.cpu().contiguous().view(torch.uint8), p.W_q.contiguous().view(torch.uint8))
    for M in (1, 2, 3, 4, 7, 300):
        x = xgen(M, K, dtype, 300 + M)
        y = layer(x.to(DEV))
        # M <= 4: streaming CUDA-core kernel (gl_gemv_mx.cu); above: CTA-pair tcgen05 kernel
        assert _cabi.last_kernel() == (MXV_NAME if M <= 4 else "gemm_a16w4_mxfp4_tcgen05_2cta"), (M, _cabi.last_kernel())
        assert y.dtype == dtype and relerr(y, O.forward(p, x)) <= REL[dtype], (dtype, M)
        if M <= 4:   # forcing "GEMM" keeps the tensor-core kernel; both kernels must agree with the oracle
            yg = layer.forward_manual(x.to(DEV), "GEMM")
            assert _cabi.last_kernel() == "gemm_a16w4_mxfp4_tcgen05_2cta" and relerr(yg, O.forward(p, x)) <= REL[dtype]
    # x = identity: every output is one dequantised weight -> bit-exact against the oracle's table * 2^(e-127)
    layer_nb = A16W4_MXFP(device=DEV, dtype=dtype).from_weights(codes.to(DEV), sc.to(DEV))
    p.bias = None
    x = torch.eye(K, dtype=dtype)
    assert torch.equal(layer_nb(x.to(DEV)).cpu(), O.forward(p, x))


def test_bitnet_and_state_dict_roundtrip(tmp_path):
    N, K = 128, 512
    g = torch.Generator().manual_seed(4)
    Wt = torch.randint(-1, 2, (N, K), generator=g).half()
    layer = A16W158_INT(device=DEV, dtype=torch.float16).from_weights(Wt, torch.tensor(0.02))
    assert (layer.W_group_mode, layer.channel_scale_mode, layer.W_nbits) == (1, 1, 2)
    x = xgen(2, K, torch.float16, 6)
    yref = (x.float() @ (Wt.float() * 0.02).t())
    assert relerr(layer(x.to(DEV)), yref) <= 3e-3

    W_q, scales, zeros = gen_int(N, K, 4, 64, torch.float16, seed=21)
    l1 = GemLiteLinear(4, 64, K, N, DType.FP16, DType.FP16).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV), None)
    keys = set(l1.state_dict().keys())
    assert keys == {"W_q", "scales", "zeros", "metadata", "orig_shape"}      # core.py:502-517 (bias absent)
    f = tmp_path / "layer.pt"
    torch.save(l1.state_dict(), f)
    l2 = GemLiteLinear()
    l2.load_state_dict(torch.load(f))
    assert l1.get_meta_args() == l2.get_meta_args()
    for a, b in zip(l1.get_tensor_args(), l2.get_tensor_args()):
        assert torch.equal(a, b)
    for M in (1, 4):
        xx = xgen(M, K, torch.float16, M).to(DEV)
        assert (l1.forward_manual(xx, "GEMM") - l2.forward_manual(xx, "GEMM")).abs().mean().item() < 1e-7


def test_error_behaviour():
    layer = GemLiteLinear(4, 64, 256, 64, DType.FP16, DType.FP16)
    W_q, scales, zeros = gen_int(64, 256, 4, 64, torch.float16, seed=1)
    layer.pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV))
    with pytest.raises(TypeError):
        layer(torch.zeros(1, 256, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(KeyError):
        layer.forward_manual(torch.zeros(1, 256, dtype=torch.float16, device=DEV), "NOPE")
    with pytest.raises(Exception, match="INT8 inputs is not compatible"):
        GemLiteLinear(4, 64, 256, 64, DType.INT8, DType.FP16).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV) + 0.5)
    with pytest.raises(AssertionError):
        B.pack_weights_over_cols(W_q.to(DEV), 3, 32, True)
    assert layer(torch.zeros(0, 256, dtype=torch.float16, device=DEV)).shape == (0, 64)


def test_split_k_under_graph_capture_without_warmup_on_the_capture_stream():
    """The split-K workspace must be zero before its first use.  When the first split-K call on a stream happens INSIDE a
    CUDA-graph capture (torch.cuda.graph's own capture stream, warm-up done elsewhere), the buffer is zeroed by a memset
    node of that graph, i.e. only when that graph replays -- so every capture gets its own buffer.  Two graphs are
    captured back to back and the SECOND is replayed first; both must reproduce the eager result bit for bit.
    (Kept last in the file: it is the only test of the capture-id path.)"""
    dtype, gs, K, N = torch.float16, 128, 1024, 2048
    W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=77)
    layer = GemLiteLinear(4, gs, K, N, DType.FP16, DType.FP16).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV))
    xs = [xgen(M, K, dtype, 880 + M).to(DEV) for M in (32, 100)]
    side = torch.cuda.Stream(device=DEV)
    with torch.cuda.stream(side):                 # warm-up and eager reference on ANOTHER stream
        want = [layer.forward_manual(x, "GEMM") for x in xs]
        assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta"
    torch.cuda.synchronize()
    graphs, outs = [], []
    for x in xs:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):                 # torch's default capture stream: no eager workspace exists for it
            outs.append(layer.forward_manual(x, "GEMM"))
        graphs.append(g)
    for i in (1, 0, 1, 0):
        graphs[i].replay()
        torch.cuda.synchronize()
        assert torch.equal(outs[i], want[i]), i


# ------------------------------------------------------------------------------------------------
def test_chained_decode_graph_matches_eager():
    """A chain of M = 1 layers (y of layer i is x of layer i + 1) captured in one CUDA graph with programmatic dependent
    launch (weights requested before griddepcontrol.wait, x after it): the result must equal the eager chain bit for bit
    on every replay, also with a foreign in-place kernel between two layers."""
    K = N = 4096
    dt = torch.float16
    layers = []
    for i in range(6):
        W_q, s, z = gen_int(N, K, 4, 128, dt, seed=900 + i)
        s = (s.float() * 0.3).to(dt)           # keep a 6-deep chain in range
        layers.append(GemLiteLinear(4, 128, K, N, DType.FP16, DType.FP16).pack(W_q.to(DEV), s.to(DEV), z.to(DEV)))
    x = xgen(1, K, dt, 77).to(DEV)

    def chain(h, act_after=None):
        for i, L in enumerate(layers):
            h = L(h)
            if act_after is not None and i == act_after:
                h.mul_(0.5)                    # foreign in-place kernel on y
        return h

    y_ref, y_ref_act = chain(x).clone(), chain(x, 2).clone()
    torch.cuda.synchronize()
    assert _cabi.last_kernel() == "gemv_w4_mma"
    for act_after, want in ((None, y_ref), (2, y_ref_act)):
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            chain(x, act_after)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                y = chain(x, act_after)
            for _ in range(10):
                g.replay()
                torch.cuda.synchronize()
                assert torch.equal(y, want), (act_after, (y != want).sum().item())
        torch.cuda.current_stream().wait_stream(st)


# ------------------------------------------------------------------------------------------------
def test_column_shards_reproduce_the_unsharded_layer_on_the_gpu():
    """SURVEY 8e on the real CUDA path: every column shard of a packed layer (ColumnShardedLinear.local), run on its own
    and concatenated in rank order -- what the all-gather produces -- equals the unsharded layer's output BIT FOR BIT
    (a column of y depends only on that column of W / scales / zeros).  A16W4 at decode and prefill batch sizes, A8W8
    int8, MXFP4 including the K/32 == N shape whose scales are square (sliced along the wrong axis)."""
    from gemlite_b200.sharding import ColumnShardedLinear
    cases = []
    W_q, s, z = gen_int(1024, 2048, 4, 128, torch.float16, seed=321)
    cases.append(("a16w4", GemLiteLinear(4, 128, 2048, 1024, DType.FP16, DType.FP16).pack(W_q.to(DEV), s.to(DEV), z.to(DEV)),
                  torch.float16, (1, 5, 16, 200)))
    Wf = (torch.randn(512, 1024, generator=torch.Generator().manual_seed(5)) / 10).half()
    cases.append(("a8w8_int8", A8W8_INT8_dynamic(device=DEV, dtype=torch.float16).from_weights(Wf.to(DEV)), torch.float16, (1, 16, 150)))
    for (n, k) in ((512, 1024), (256, 8192)):                      # 8192 / 32 == 256: square MX scale matrix
        lin = torch.nn.Linear(k, n, bias=False, dtype=torch.bfloat16, device=DEV)
        lin.weight.data = (torch.randn(n, k, generator=torch.Generator().manual_seed(n)) / 10).to(device=DEV, dtype=torch.bfloat16)
        cases.append((f"mxfp4_{n}x{k}", A16W4_MXFP(device=DEV, dtype=torch.bfloat16).from_linear(lin, del_orig=False), torch.bfloat16, (1, 4, 200)))
    for name, layer, dt, Ms in cases:
        for world in (2, 4):
            shards = [ColumnShardedLinear(layer, r, world).local for r in range(world)]
            for M in Ms:
                x = xgen(M, layer.in_features, dt, 40 + M).to(DEV)
                full = layer(x)
                k_full = _cabi.last_kernel()
                parts = torch.cat([sh(x) for sh in shards], dim=-1)
                k_shard = _cabi.last_kernel()
                if k_shard == k_full:          # same kernel, same per-column summation order: bit-equal
                    assert torch.equal(parts, full), (name, world, M, k_full, (parts != full).sum().item())
                else:                          # a narrow shard can fall to another kernel (tile-width rules): fp tolerance
                    assert relerr(parts, full) <= REL[dt], (name, world, M, k_full, k_shard)


# ------------------------------------------------------------------------------------------------
# Batched decode, 16 < M <= 128: gl_gemm_mid.cu (cluster of K/1024 CTAs per 128-column slab, mma.sync, exact dequant)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(256, 1024), (512, 2048), (384, 4096), (128, 8192)])
def test_a16w4_batched_decode_kernel_vs_oracle(dtype, shape):
    N, K = shape
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    for gs, mode in ((128, "fma"), (64, "nofma"), (256, "sym"), (64, "fma")):
        W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=gs + K // 64 + (3 if mode == "fma" else 5))
        bias = (torch.randn(N, generator=torch.Generator().manual_seed(6)) / 10).to(dtype) if mode != "sym" else None
        z_in = None if mode == "sym" else zeros
        layer = GemLiteLinear(4, gs, K, N, gd, gd)
        layer.pack(W_q.to(DEV), scales.to(DEV), None if z_in is None else z_in.to(DEV), None if bias is None else bias.to(DEV),
                   fma_mode=(mode == "fma"))
        p = O.prepare(W_q, scales, z_in, W_nbits=4, group_size=gs, in_features=K, out_features=N,
                      input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], bias=bias, fma_mode=(mode == "fma"))
        for M in (17, 32, 63, 64, 65, 100, 128):          # one and two token passes, ragged tails
            x = xgen(M, K, dtype, 300 + M)
            y = layer(x.to(DEV))
            assert _cabi.last_kernel() == "gemm_a16w4_mma_mid", (M, _cabi.last_kernel())
            e = relerr(y, O.forward(p, x))
            assert e <= REL[dtype], (dtype, shape, gs, mode, M, e)
            assert torch.equal(layer(x.to(DEV)), y)       # deterministic: fixed-order reductions, no atomics
        # the explicitly named kernels of the reference still resolve: GEMM -> tcgen05 pair kernel, GEMM_SPLITK -> this one
        x = xgen(40, K, dtype, 9).to(DEV)
        layer.forward_manual(x, "GEMM_SPLITK")
        assert _cabi.last_kernel() == "gemm_a16w4_mma_mid"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_batched_decode_kernel_dequant_is_bit_exact(dtype):
    """One-hot activations (row m of x = e_{k0 + m}) => y[m, n] is the dequantised weight W[n, k0 + m] itself: the MMA
    operand must hold exactly the reference's dequantize() values (utils.py:57-89), for every k position of a word and
    both halves of the CTA's k range."""
    N, K, gs = 256, 2048, 64
    W_q, scales, zeros = gen_int(N, K, 4, gs, dtype, seed=15)
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    layer = GemLiteLinear(4, gs, K, N, gd, gd).pack(W_q.to(DEV), scales.to(DEV), zeros.to(DEV))
    p = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype])
    for k0 in (0, 960, 1024 + 37, K - 128):
        x = torch.zeros(128, K, dtype=dtype)
        x[torch.arange(128), k0 + torch.arange(128)] = 1
        y = layer(x.to(DEV))
        assert _cabi.last_kernel() == "gemm_a16w4_mma_mid"
        assert torch.equal(y.cpu(), O.forward(p, x)), k0


@pytest.mark.parametrize("kind", ["int8", "fp8"])
@pytest.mark.parametrize("shape", [(256, 1024), (384, 2048), (256, 4096), (128, 8192)])
def test_a8w8_batched_decode_kernel(kind, shape):
    """gl_gemm_a8_mid.cu, 16 < M <= 128 (one and two token passes, ragged tails), clusters of 1 / 2 / 4 / 8 CTAs: int8 must
    equal the exact integer product + the one-rounding epilogue bit for bit on the product's own quantised activations
    (and therefore the tcgen05 kernels and the reference, test_gpu_vs_reference.py); fp8 within 2e-3 of the fp64 product
    of the same codes.  With and without bias; deterministic."""
    N, K = shape
    W = (torch.randn(N, K, generator=torch.Generator().manual_seed(K // 8 + N)) / 10).half()
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(2)) / 10).half()
    cls = A8W8_INT8_dynamic if kind == "int8" else A8W8_FP8_dynamic
    wdt = torch.int8 if kind == "int8" else torch.float8_e4m3fn
    for b in (None, bias):
        layer = cls(device=DEV, dtype=torch.float16).from_weights(W.clone().to(DEV), bias=None if b is None else b.to(DEV))
        Wc = layer.W_q.data.t().contiguous()                          # (N, K) codes
        sw = layer.scales.data.view(1, -1).float()
        for M in (17, 32, 40, 64, 65, 100, 128):
            x = xgen(M, K, torch.float16, 70 + M).to(DEV)
            y = layer(x)
            assert _cabi.last_kernel() == f"gemm_a8w8_{kind}_mma_mid", (M, _cabi.last_kernel())
            q, s = scale_activations_per_token(x, w_dtype=wdt)
            if kind == "int8":
                acc = (q.to(torch.int32).cpu() @ Wc.to(torch.int32).cpu().t()).float().to(DEV)
                ref = (acc * (s.float().view(-1, 1) * sw)).half()
                if b is not None:
                    ref = (ref.float() + b.to(DEV).float()).half()
                assert torch.equal(y, ref), (shape, M, b is not None, (y != ref).sum().item())
            else:
                ref = (q.double() @ Wc.double().t()) * (s.double().view(-1, 1) * sw.double())
                if b is not None:
                    ref = ref + b.to(DEV).double()
                assert relerr(y, ref) <= 2e-3, (shape, M, relerr(y, ref))
            assert torch.equal(layer(x), y)


@pytest.mark.parametrize("kind", ["int8", "fp8"])
@pytest.mark.parametrize("post_scale", [False, True])
def test_a16w8_weight_only_decode_kernel(kind, post_scale):
    """helper.A16W8_INT8 / A16W8_FP8 at M <= 16 on the streaming kernel of gl_gemv_a8.cu (fp16 activations, weights
    widened to fp16 in registers, per-channel scale on the fp32 accumulator) against the oracle; 16 < M <= 128 runs on
    the clustered batched-decode kernel of gl_gemm_a8_mid.cu (same arithmetic).  Pre-scale mode differs from the oracle by the fp16 rounding of every q * s (decode tolerance)."""
    N, K = 512, 2048
    W = (torch.randn(N, K, generator=torch.Generator().manual_seed(77)) / 10).half()
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(3)) / 10).half()
    cls = A16W8_INT8 if kind == "int8" else A16W8_FP8
    for b in (None, bias):
        layer = cls(device=DEV, dtype=torch.float16, post_scale=post_scale).from_weights(W.clone().to(DEV), bias=None if b is None else b.to(DEV))
        p = oracle_from_layer(layer)
        for M in (1, 3, 8, 16, 40):
            x = xgen(M, K, torch.float16, 60 + M)
            y = layer(x.to(DEV))
            want = f"gemv_a16w8_{kind}" if M <= 16 else f"gemm_a16w8_{kind}_mma_mid"
            assert _cabi.last_kernel() == want, (M, _cabi.last_kernel())
            e = relerr(y, O.forward(p, x))
            assert e <= REL_DECODE[torch.float16], (kind, post_scale, b is not None, M, e)


@pytest.mark.parametrize("kind", ["int8", "fp8"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("post_scale", [False, True])
def test_a16w8_weight_only_pair_kernel(kind, dtype, post_scale):
    """helper.A16W8_INT8 / A16W8_FP8 above the decode range (and bf16 at every M) on the CTA-pair tcgen05 kernel: the
    8-bit weights are widened by the dequant warps straight into tensor memory (pre-scaling: round16(float(q) * s), the
    reference's rounding point; post-scaling: exact values, accumulator x s in the epilogue) -- against the oracle,
    split-K range included, and bit for bit through one-hot activations."""
    N, K = 768, 2048
    W = (torch.randn(N, K, generator=torch.Generator().manual_seed(78)) / 10).to(dtype)
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(4)) / 10).to(dtype)
    cls = A16W8_INT8 if kind == "int8" else A16W8_FP8
    for b in (None, bias):
        layer = cls(device=DEV, dtype=dtype, post_scale=post_scale).from_weights(W.clone().to(DEV), bias=None if b is None else b.to(DEV))
        p = oracle_from_layer(layer)
        for M in (17, 64, 129, 300, 1000):
            if dtype == torch.bfloat16 or M > 16:
                x = xgen(M, K, dtype, 90 + M)
                y = layer(x.to(DEV))
                mid = 16 < M <= 128                    # batched-decode kernel (pre-scale rounding reproduced since visit 22)
                assert _cabi.last_kernel() == (f"gemm_a16w8_{kind}_mma_mid" if mid else f"gemm_a16w8_{kind}_tcgen05_2cta"), (M, _cabi.last_kernel())
                e = relerr(y, O.forward(p, x))
                assert e <= REL[dtype], (kind, dtype, post_scale, b is not None, M, e)
                assert torch.equal(layer(x.to(DEV)), y)
                if mid:                                # the tcgen05 kernel by name, same layer
                    y2 = layer.forward_manual(x.to(DEV), "GEMM")
                    assert _cabi.last_kernel() == f"gemm_a16w8_{kind}_tcgen05_2cta"
                    assert relerr(y2, O.forward(p, x)) <= REL[dtype]
    if dtype == torch.bfloat16:
        x1 = xgen(1, K, dtype, 5)
        y1 = layer(x1.to(DEV))
        assert _cabi.last_kernel() == f"gemm_a16w8_{kind}_tcgen05_2cta"
        assert relerr(y1, O.forward(p, x1)) <= REL[dtype]
    # one-hot activations: every output is one dequantised weight (times 1, plus bias) -> the oracle's bits
    layer = cls(device=DEV, dtype=dtype, post_scale=post_scale).from_weights(W.clone().to(DEV))
    p = oracle_from_layer(layer)
    eye = torch.eye(K, dtype=dtype)[:512]
    y = layer(eye.to(DEV)).cpu()
    ref = O.forward(p, eye)
    assert torch.equal(y, ref.to(y.dtype)), (kind, dtype, post_scale, (y != ref.to(y.dtype)).sum().item())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("nbits", [8, 2, 1])
@pytest.mark.parametrize("mode", ["fma", "nofma", "sym"])
def test_packed_8_2_1_bit_layers_on_the_pair_kernel(dtype, nbits, mode):
    """A16W8 / A16W2 / A16W1 grouped layers (packing_bitwidth 32; helper.py:344-400) take the same TMA -> unpack ->
    tensor-memory -> tcgen05 path as 4-bit ones above the decode range (and at every M by name): against the oracle, and
    through one-hot activations bit for bit (the tensor-memory operand holds exactly the reference's dequantize() values);
    M <= 8 runs the streaming kernel of gl_gemv_packed.cu."""
    N, K, gs = 512, 1024, 64
    W_q, scales, zeros = gen_int(N, K, nbits, gs, dtype, seed=nbits + (31 if mode == "fma" else 37))
    gd = gemlite_b200.dtypes.TORCH_TO_DTYPE[dtype]
    bias = (torch.randn(N, generator=torch.Generator().manual_seed(4)) / 10).to(dtype)
    z_in = None if mode == "sym" else zeros
    name = {8: "gemm_a16w8_packed_tcgen05_2cta", 2: "gemm_a16w2_tcgen05_2cta", 1: "gemm_a16w1_tcgen05_2cta"}[nbits]
    for b in (None, bias):
        layer = GemLiteLinear(nbits, gs, K, N, gd, gd)
        layer.pack(W_q.to(DEV), scales.to(DEV), None if z_in is None else z_in.to(DEV), None if b is None else b.to(DEV),
                   fma_mode=(mode == "fma"), packing_bitwidth=32)
        p = O.prepare(W_q, scales, z_in, W_nbits=nbits, group_size=gs, in_features=K, out_features=N,
                      input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], bias=b, fma_mode=(mode == "fma"),
                      packing_bitwidth=32)
        for M in (1, 3, 8, 16, 64, 129, 300):
            x = xgen(M, K, dtype, 90 + M)
            y = layer(x.to(DEV))
            # M <= 8: streaming CUDA-core kernel (gl_gemv_packed.cu; scale / zero factorised per group: decode tolerance);
            # the tensor-core kernel stays selectable by name
            small = M <= 8
            assert _cabi.last_kernel() == (f"gemv_w{nbits}_packed" if small else name), (M, _cabi.last_kernel())
            e = relerr(y, O.forward(p, x))
            assert e <= (REL_DECODE if small else REL)[dtype], (dtype, nbits, mode, b is not None, M, e)
            assert torch.equal(layer(x.to(DEV)), y)
            if small:
                y2 = layer.forward_manual(x.to(DEV), "GEMM")
                assert _cabi.last_kernel() == name
                assert relerr(y2, O.forward(p, x)) <= REL[dtype]
    layer = GemLiteLinear(nbits, gs, K, N, gd, gd)
    layer.pack(W_q.to(DEV), scales.to(DEV), None if z_in is None else z_in.to(DEV), None, fma_mode=(mode == "fma"), packing_bitwidth=32)
    p = O.prepare(W_q, scales, z_in, W_nbits=nbits, group_size=gs, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[dtype], output_dtype=O.DTYPE_OF[dtype], fma_mode=(mode == "fma"), packing_bitwidth=32)
    eye = torch.eye(K, dtype=dtype)
    assert torch.equal(layer(eye.to(DEV)).cpu(), O.forward(p, eye))
