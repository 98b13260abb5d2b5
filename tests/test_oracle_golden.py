"""CPU: the oracle (oracle/gemlite_oracle.py) against fixtures produced by the reference itself.

Every fixture in tests/golden/ was written by tests/golden/make_golden.py, which executes the
unmodified reference (pack(), its Triton kernels under the CPU interpreter, its quantisers).
"""
import numpy as np
import pytest
import torch

from oracle import gemlite_oracle as O
from oracle import recipes as R
from golden_io import MANIFEST, MX_DYNAMIC, from_bits, layer_cases, load, tdtype

# fp32-accumulating reference kernels agree with the oracle to an output ulp or two; the reference's
# GEMV kernels accumulate in the input dtype (fp16/bf16) and are compared with a wider bound.
TOL_REL_ACC32 = {torch.float16: 2.0e-3, torch.bfloat16: 1.6e-2, torch.float32: 1e-5}
TOL_REL_GEMV = {torch.float16: 2.5e-2, torch.bfloat16: 1.5e-1, torch.float32: 1e-4}


def test_bitpack_matches_reference():
    z, info = load("bitpack")
    for nbits in (8, 4, 2, 1):
        W_q = z[f"in_b{nbits}"]
        for pw in (8, 16, 32, 64):
            if pw < nbits:
                continue
            for tr in (1, 0):
                out, eps = O.pack_over_cols(W_q, nbits, pw, bool(tr))
                ref = z[f"torch_b{nbits}_w{pw}_t{tr}"]
                assert eps == pw // nbits
                key = f"triton_b{nbits}_w{pw}_t{tr}"
                if pw == 64:
                    # reference quirk: the *torch* packer shifts int32 values by up to 56 bits
                    # (bitpack.py:44,54), so its 64-bit words are wrong; the Triton packer -- the one
                    # pack() uses on CUDA (core.py:385) -- is the contract there.
                    assert key in z.files
                else:
                    assert out.shape == ref.shape and np.array_equal(np.ascontiguousarray(out), ref), (nbits, pw, tr)
                if key in z.files:
                    assert np.array_equal(np.ascontiguousarray(out), z[key]), key
                back = O.unpack_over_cols(out, nbits, pw, bool(tr))
                assert np.array_equal(back, W_q)


def _interpreter_untrustworthy(name, p):
    """Triton's CPU interpreter keeps bf16 tensors as raw uint16 and does ARITHMETIC on the bits, so kernels that
    compute in bf16 (the INT-weight GEMV / GEMM kernels with bf16 activations) produce garbage there; their forward
    fixtures are not reference behaviour on a GPU and are not compared.  (The MX kernels only touch bf16 in loads /
    the final store, whose conversions make_golden.py corrects, and fp8 fixtures are exact since that correction.)"""
    return ("bf16" in name) and p.input_dtype not in O.MX_VALUES


def build_oracle_layer(name, z, info):
    """Re-create the layer with the oracle from the *inputs* stored in the fixture."""
    xdt = tdtype(info["x_dtype"])
    bias = None
    if name in MX_DYNAMIC:
        kind, post = MX_DYNAMIC[name]
        W = from_bits(z["in_W"], torch.bfloat16)
        p, codes, sc = R.mx_dynamic(W, kind, torch.bfloat16, post_scale=post)
        # the reference's weight quantisers (mxfp8 / mxfp4 / nvfp4) must be reproduced bit for bit
        assert np.array_equal(codes.view(torch.uint8).numpy(), z["q_codes"]), "MX codes differ from reference"
        assert np.array_equal(sc.view(torch.uint8).numpy(), z["q_scales"]), "MX scales differ from reference"
        return p
    if name.startswith("a16w") and "in_W_q" in z.files:
        parts = name.split("_")
        nbits = int(parts[0][4:])
        dtype = torch.bfloat16 if "bf16" in name else torch.float16
        gs = 256 if "channelwise" in name else int([p for p in parts if p.startswith("gs")][0][2:])
        W_q = torch.from_numpy(z["in_W_q"])
        scales = from_bits(z["in_scales"], dtype)
        zeros = from_bits(z["in_zeros"], dtype)
        if "in_bias" in z.files:
            bias = from_bits(z["in_bias"], dtype)
        pw = 8 if name.endswith("pw8") else 32
        N, K = W_q.shape
        gd = O.DTYPE_OF[dtype]   # make_golden.py calls pack() directly (no helper mode override)
        return O.prepare(W_q, scales, zeros, W_nbits=nbits, group_size=gs, in_features=K, out_features=N,
                         input_dtype=gd, output_dtype=gd, bias=bias, fma_mode="nofma" not in name,
                         packing_bitwidth=pw)
    if name.startswith("a16w8_int8"):
        W = from_bits(z["in_W"], torch.float16)
        return R.a16w8(W, torch.float16, post_scale=name.endswith("postscale"))
    if name == "int8_w4_scalar_zero":
        W_q = torch.from_numpy(z["in_W_q"])
        sc = from_bits(z["in_scales"], torch.float16)
        N, K = W_q.shape
        return O.prepare(W_q, sc, 7, W_nbits=4, group_size=K, in_features=K, out_features=N,
                         input_dtype=O.INT8, output_dtype=O.FP32, scaled_activations=False)
    if name.startswith("a8w8_"):
        W = from_bits(z["in_W"], torch.float16)
        return R.a8w8_dynamic(W, torch.float16, fp8=torch.float8_e4m3fn if "fp8" in name else None)
    if "mxfp" in name:
        W = from_bits(z["in_W"], torch.bfloat16)
        nbits = 4 if name.startswith("a16w4") else 8
        p, codes, sc = R.a16wn_mxfp(W, nbits, torch.bfloat16)
        # the reference's own quantiser output must be reproduced bit for bit
        assert np.array_equal(codes.view(torch.uint8).numpy(), z["q_codes"]), "MX codes differ from reference"
        assert np.array_equal(sc.view(torch.uint8).numpy(), z["q_scales"]), "MX scales differ from reference"
        return p
    raise KeyError(name)


@pytest.mark.parametrize("name", layer_cases())
def test_pack_and_forward_match_reference(name):
    z, info = load(name)
    p = build_oracle_layer(name, z, info)
    # ---- pack(): layout, mode resolution and metadata are bit-exact -------------------------------
    assert p.meta_args() == [int(v) for v in z["meta_args"]], (p.meta_args(), z["meta_args"])
    assert list(p.W_q.shape) == info["W_q_shape"]
    assert list(p.W_q.stride()) == [int(v) for v in z["packed_W_q_strides"]]
    wq = p.W_q.contiguous()
    wq_bits = wq.view(torch.uint8).numpy() if wq.dtype.is_floating_point else wq.numpy()
    assert np.array_equal(wq_bits, z["packed_W_q"]), "packed W_q differs"
    sdt, zdt = tdtype(info["scales_dtype"]), tdtype(info["zeros_dtype"])
    assert p.scales.dtype == sdt and p.zeros.dtype == zdt
    assert list(p.scales.shape) == info["scales_shape"] and list(p.zeros.shape) == info["zeros_shape"]
    assert torch.equal(p.scales.contiguous(), from_bits(z["packed_scales"], sdt).view(p.scales.shape))
    assert torch.equal(p.zeros.contiguous(), from_bits(z["packed_zeros"], zdt).view(p.zeros.shape))

    # ---- forward(): every reference kernel the interpreter could execute ----------------------------
    xdt = tdtype(info["x_dtype"])
    out_t = O.TORCH_OF[p.output_dtype]
    if _interpreter_untrustworthy(name, p):
        return  # pack()/quantiser parts above are still pinned; see the helper for why forward is not
    checked = 0
    i = 0
    while f"x{i}" in z.files:
        x = from_bits(z[f"x{i}"], xdt)
        M = x.shape[0]
        y = O.forward(p, x).to(torch.float64)
        scale = y.abs().max().item() + 1e-30
        for mt in ("GEMV_REVSPLITK", "GEMV", "GEMV_SPLITK", "GEMM_SPLITK", "GEMM"):
            key = f"y_{mt}_M{M}"
            if key not in z.files:
                continue
            if mt == "GEMV" and p.input_dtype in O.MX_VALUES:
                continue  # the reference documents its MX GEMV as broken (core.py:103 "outputs garbage"): inf / 1e34 here
            if mt == "GEMV" and p.channel_scale_mode == 1:
                continue  # reference quirk: plain GEMV drops channel scaling when scales_x is None (gemv_kernels.py:593-595)
            if mt == "GEMV_SPLITK" and p.W_group_mode > 0 and p.group_size < p.in_features:
                continue  # reference default config BLOCK_SIZE_K=2048 > group_size reads one scale row per block
                          # (gemv_splitK_kernels.py:152,326): not a valid oracle for grouped weights
            if mt == "GEMV_REVSPLITK" and p.elements_per_sample > 16:
                continue  # reference default BLOCK_SIZE_K=16 < elements_per_sample=32 (gemv_revsplitK_kernels.py:145,340)
            ref = from_bits(z[key], out_t).to(torch.float64)
            err = (y - ref).abs().max().item() / scale
            if mt == "GEMV" and p.group_size == p.in_features and p.W_group_mode > 0:
                continue  # manual-only kernel (core.py:102-105); with one group spanning K its default config
                          # indexes scale rows past the single row -- not a usable oracle
            if mt.startswith("GEMM"):
                tol = TOL_REL_ACC32[out_t]
            elif xdt == torch.int8:
                tol = 2.0e-3   # exact int32 partial sums, but scaled per K-chunk in the fp16 meta dtype before the atomic add
            else:
                tol = TOL_REL_GEMV[out_t]
            assert err <= tol, (name, key, err, tol)
            if name in EXACT_A8 and mt in ("GEMM", "GEMM_SPLITK", "GEMV_SPLITK"):
                # 8-bit activations x 8-bit weights: the int32 (and, on these sizes, the fp32-of-fp8) accumulator is exact
                # and the epilogue is one rounding -- the oracle must reproduce the reference's output BITS, so that a
                # future oracle change inside the float tolerance cannot silently void "bit-exact" 
                assert torch.equal(y, ref), (name, key, err)
            if p.input_dtype in O.MX_VALUES and not (p.scaled_activations and p.channel_scale_mode == 2 and mt == "GEMM_SPLITK"):
                # MX paths: quantisers, scale indexing, NVFP4 norm and the output rounding are reproduced bit for bit
                # (the post-scale split-K kernel scales in its fp32 buffer, one rounding fewer than the GEMM kernel
                # the oracle follows: gemm_splitK_kernels.py:577-581,606-607 vs gemm_kernels.py:533-537)
                assert torch.equal(y, ref), (name, key, err)
            checked += 1
        i += 1
    assert checked > 0


EXACT_A8 = ("a8w8_int8_dynamic", "a8w8_fp8_dynamic", "int8_w4_scalar_zero")


def test_activation_quant_matches_reference():
    for name, wdt in (("a8w8_int8_dynamic", torch.int8), ("a8w8_fp8_dynamic", torch.float8_e4m3fn)):
        z, info = load(name)
        i = 0
        seen = 0
        while f"x{i}" in z.files and f"actq_q{i}" in z.files:
            x = from_bits(z[f"x{i}"], torch.float16)
            q, s = O.quantize_activations_per_token(x, wdt)
            qr = from_bits(z[f"actq_q{i}"], wdt)
            sr = torch.from_numpy(z[f"actq_s{i}"])
            assert torch.equal(s, sr.view(s.shape)), name
            assert torch.equal(q.view(torch.uint8), qr.view(torch.uint8).view(q.shape)), name
            seen += 1
            i += 1
        assert seen > 0, name


def test_mx_activation_quantisers_match_reference():
    """scale_activations_mxfp8 / mxfp4 / nvfp4 (the reference's *_triton_v2 kernels under the interpreter): codes AND
    scales bit for bit, including the planted edge rows (all-zero group, exact-power-of-two maxima with elements on the
    e2m1 rounding thresholds, 2^-30 exponent clamp, large values) and the scale rows of the padding up to 32 / 16 rows."""
    z, info = load("actq_mx")
    fns = {"mxfp8": O.quantize_activations_mxfp8, "mxfp4": O.quantize_activations_mxfp4, "nvfp4": O.quantize_activations_nvfp4}
    seen = 0
    for tag, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        for M in (1, 5, 33):
            x = from_bits(z[f"x_{tag}_M{M}"], dt)
            for kind, fn in fns.items():
                key = f"{kind}_{tag}_M{M}"
                assert key in info["kernels_ok"], key
                q, s = fn(x)
                qr, sr = z["q_" + key], z["s_" + key]
                assert q.shape == qr.shape and s.shape == sr.shape, (key, q.shape, qr.shape, s.shape, sr.shape)
                assert np.array_equal(q.view(torch.uint8).numpy(), qr), key
                assert np.array_equal(s.view(torch.uint8).numpy(), sr), key
                seen += 1
    assert seen == 18


def test_c_cpu_baseline_matches_the_oracle():
    """oracle/cpu_a16w4.c (gcc + OpenMP; bench.py's cpu_baseline and --impl reference legs) against the torch port and the
    rounding-exact oracle on a seeded A16W4 layer: fp32 summation-order noise only."""
    import numpy as np
    from oracle import build_cpu as C
    N, K, gs = 256, 1024, 128
    g = torch.Generator().manual_seed(5)
    W_q = torch.randint(0, 16, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    scales = (torch.rand(N * K // gs, 1, generator=g) * 0.01 + 0.001).half()
    zeros = (torch.rand(N * K // gs, 1, generator=g) * 15).half()
    p = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=gs, in_features=K, out_features=N)
    assert p.W_group_mode == 4
    x = (torch.randn(3, K, generator=g) / 10).half()
    y_c = C.forward(p.W_q.contiguous().numpy(), p.scales.float().numpy(), p.zeros.float().numpy(), x.float().numpy())
    y_t = O.cpu_dequant_matmul_fast(p.W_q.contiguous(), p.scales, p.zeros, x, gs).numpy()
    y_o = O.forward(p, x).float().numpy()
    scale = np.abs(y_o).max()
    assert np.abs(y_c - y_t).max() <= 1e-5 * scale
    assert np.abs(y_c - y_o).max() <= 2e-3 * scale       # the oracle rounds every dequantised weight and the output to fp16
