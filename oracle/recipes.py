"""Quantise-and-pack recipes of the reference's helper processors, restated on top of the oracle.

TEST INFRASTRUCTURE (same rule as gemlite_oracle.py).  Each function cites the helper it follows.
"""
import torch

from . import gemlite_oracle as O


def _per_channel_quant(W, w_dtype):
    """helper.py:104-128 / 425-444: fp32 absmax per output channel, clamp 1e-6, round-half-even."""
    Wf = W.to(torch.float32)
    if w_dtype.is_floating_point:
        lo, hi = torch.finfo(w_dtype).min, torch.finfo(w_dtype).max
    else:
        lo, hi = torch.iinfo(w_dtype).min, torch.iinfo(w_dtype).max
    scales = Wf.abs().amax(dim=1, keepdim=True) / hi
    scales.clamp_(min=1e-6)
    W_q = (Wf / scales).clamp_(min=lo, max=hi)
    W_q = W_q.to(w_dtype) if w_dtype.is_floating_point else W_q.round_().to(w_dtype)
    return W_q, scales


def a16w8(W, dtype, fp8=None, post_scale=False, bias=None):
    """A16W8 / A16W8_INT8 / A16W8_FP8.from_weights (helper.py:88-165)."""
    N, K = W.shape
    W_q, scales = _per_channel_quant(W, fp8 if fp8 else torch.int8)
    gd = O.DTYPE_OF[dtype]
    p = O.prepare(W_q, scales, None, W_nbits=8, group_size=K, in_features=K, out_features=N,
                  input_dtype=gd, output_dtype=gd, bias=bias)
    p.W_group_mode, p.channel_scale_mode = (0, 1) if post_scale else (2, 0)
    return p


def a8w8_dynamic(W, dtype, fp8=None, bias=None):
    """A8W8_dynamic.from_weights (helper.py:405-476): both operands 8-bit, scales applied post-accumulate."""
    N, K = W.shape
    w_dtype = fp8 if fp8 else torch.int8
    W_q, scales = _per_channel_quant(W, w_dtype)
    p = O.prepare(W_q, scales.to(torch.float32), None, W_nbits=8, group_size=K, in_features=K, out_features=N,
                  input_dtype=O.DTYPE_OF[w_dtype], output_dtype=O.DTYPE_OF[dtype],
                  scaled_activations=True, bias=bias)
    p.W_group_mode, p.channel_scale_mode = 0, 3
    return p


def a16wn_int(W_q, scales, zeros, nbits, group_size, dtype, bias=None, fma_mode=True, packing_bitwidth=32,
              post_scale=False):
    """A16Wn.from_weights_ quant_type='INT' (helper.py:191-263)."""
    N, K = W_q.shape
    gd = O.DTYPE_OF[dtype]
    p = O.prepare(W_q, scales.to(dtype), zeros.to(dtype), W_nbits=nbits, group_size=group_size, in_features=K,
                  out_features=N, input_dtype=gd, output_dtype=gd, bias=bias, fma_mode=fma_mode,
                  packing_bitwidth=packing_bitwidth)
    if group_size == K:
        p.W_group_mode, p.channel_scale_mode = (1, 1) if post_scale else (3, 0)
    return p


def a16wn_mxfp(W, nbits, dtype, bias=None):
    """A16Wn_MXFP.from_linear -> from_weights_ quant_type='MXFP' (helper.py:299-328, 217-233)."""
    N, K = W.shape
    if nbits == 4:
        codes, sc = O.quantize_mxfp4(W)
    else:
        codes, sc = O.quantize_mxfp8(W)
    codes, sc = codes.view(N, K), sc.view(N, K // 32)
    gd = O.MXFP16 if dtype == torch.float16 else O.MXBF16
    p = O.prepare(codes, sc, None, W_nbits=nbits, group_size=32, in_features=K, out_features=N,
                  input_dtype=gd, output_dtype=gd, bias=bias)
    return p, codes, sc


def mx_dynamic(W, kind, dtype, post_scale=False, bias=None):
    """A8W8_MXFP_dynamic / A8W4_MXFP_dynamic / A4W4_MXFP_dynamic / A4W4_NVFP_dynamic .from_linear -> from_weights
    (helper.py:660-945): MX / NV quantised weights, activations quantised on the fly.  kind: 'a8w8' | 'a8w4' |
    'a4w4_mx' | 'a4w4_nv'.  Returns (packed layer, codes, scales) like a16wn_mxfp."""
    N, K = W.shape
    if kind == "a8w8":
        codes, sc = O.quantize_mxfp8(W)
        nbits, gs, in_dt = 8, 32, O.MXFP8
    elif kind == "a8w4":
        codes, sc = O.quantize_mxfp4(W)
        nbits, gs, in_dt = 4, 32, O.MXFP8
    elif kind == "a4w4_mx":
        codes, sc = O.quantize_mxfp4(W)
        nbits, gs, in_dt = 4, 32, O.MXFP4
    elif kind == "a4w4_nv":
        codes, sc = O.quantize_nvfp4(W)
        nbits, gs, in_dt = 4, 16, O.NVFP4
    else:
        raise KeyError(kind)
    codes, sc = codes.view(N, K), sc.view(N, K // gs)
    p = O.prepare(codes, sc, None, W_nbits=nbits, group_size=gs, in_features=K, out_features=N,
                  input_dtype=in_dt, output_dtype=O.DTYPE_OF[dtype], scaled_activations=True, bias=bias)
    p.W_group_mode = 0
    p.channel_scale_mode = (2 if post_scale else 4) if kind in ("a8w8", "a8w4") else 4
    return p, codes, sc
