"""CPU oracle for the GemLite hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.  Nothing under `gemlite_b200/` does.

It restates, with numpy / torch-CPU and the reference's rounding points, what the
reference computes on the path `pack() -> forward()`:

  pack_over_cols / unpack_over_cols .... gemlite/bitpack.py:36-60 (+ inverse :146-170)
  prepare() ............................ gemlite/core.py:336-519 (layout + mode resolution)
  dequantize_weights() ................. gemlite/triton_kernels/utils.py:57-89
  forward() ............................ gemlite/triton_kernels/gemm_kernels.py:347-413
                                         (scale row g = k // group_size, fp32/int32
                                         accumulate, channel-scale epilogue, cast) and
                                         gemlite/core.py:192-193 (bias)
  quantize_activations_per_token() ..... gemlite/quant_utils.py:231-305
  quantize_activations_mxfp8/mxfp4/nvfp4  gemlite/quant_utils.py:381-391,503-591,770-954
  mx_* ................................. gemlite/quant_utils.py:27-37,126-163 (e2m1 table,
                                         MXFP4/MXFP8 weight quantiser) and
                                         gemlite/triton_kernels/gemm_kernels.py:482-531
                                         (tl.dot_scaled semantics, e8m0 = 2^(e-127))

Parity status: PINNED.  `tests/golden/*.npz` hold outputs of the reference itself
(its Triton kernels executed in this container by Triton's CPU interpreter; generator:
`tests/golden/make_golden.py`) and `tests/test_oracle_golden.py` checks this module
against every one of them.

Deliberate difference from the reference, documented in DESIGN.md: the reference's
GEMV kernels accumulate in the *input* dtype (fp16) with atomics
(gemv_revsplitK_kernels.py:426-430,387); this oracle -- like the reference's own GEMM
kernels on B200 (core.py:39-54: B200 is not in gpu_supports_float16_acc) -- accumulates
in fp32, so it is the more accurate of the two and is what the CUDA kernels match.
"""
from __future__ import annotations

import numpy as np
import torch

# --- DType wire values (gemlite/dtypes.py:8-29) -------------------------------------------
FP32, FP16, BF16, FP8, INT8, UINT8, INT32 = 0, 1, 2, 3, 4, 5, 6
FP8E5 = 8
MXFP16, MXBF16, MXFP8, MXFP4, NVFP4, E8M0 = 14, 15, 16, 17, 18, 19
MX_VALUES = (MXFP16, MXBF16, MXFP8, MXFP4, NVFP4)

TORCH_OF = {
    FP32: torch.float32, FP16: torch.float16, BF16: torch.bfloat16, FP8: torch.float8_e4m3fn,
    INT8: torch.int8, UINT8: torch.uint8, INT32: torch.int32, FP8E5: torch.float8_e5m2,
    MXFP16: torch.float16, MXBF16: torch.bfloat16, MXFP8: torch.float8_e4m3fn,
    MXFP4: torch.uint8, NVFP4: torch.uint8, E8M0: torch.float8_e8m0fnu,
}
DTYPE_OF = {
    torch.float32: FP32, torch.float16: FP16, torch.bfloat16: BF16, torch.float8_e4m3fn: FP8,
    torch.int8: INT8, torch.uint8: UINT8, torch.int32: INT32, torch.float8_e5m2: FP8E5,
    torch.float8_e8m0fnu: E8M0,
}
_WORD = {8: np.uint8, 16: np.int16, 32: np.int32, 64: np.int64}
_UWORD = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}

# e2m1 code -> value; index is the 4-bit pattern, sign = bit 3 (quant_utils.py:31-37)
FP4_VALUES = np.array([0, 0.5, 1, 1.5, 2, 3, 4, 6, -0.0, -0.5, -1, -1.5, -2, -3, -4, -6], dtype=np.float32)
FP4_THRESHOLDS = np.array([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0], dtype=np.float32)


# ============================================================================================
# bit packing  (bitpack.py:36-60)
# ============================================================================================
def pack_over_cols(W_q, nbits: int, packing_bitwidth: int = 32, transpose: bool = True):
    """Element j of a row goes to word j // eps, bit offset (j % eps) * nbits."""
    assert packing_bitwidth in (8, 16, 32, 64) and nbits in (8, 4, 2, 1)
    W = np.asarray(W_q).astype(np.uint64)
    rows, cols = W.shape
    eps = packing_bitwidth // nbits
    W = W.reshape(rows, cols // eps, eps)
    shifts = (np.arange(eps, dtype=np.uint64) * np.uint64(nbits))
    words = np.bitwise_or.reduce(W << shifts, axis=2)
    out = words.astype(_UWORD[packing_bitwidth]).view(_WORD[packing_bitwidth])
    if transpose:
        out = out.T
    return out, eps


def unpack_over_cols(W_p, nbits: int, packing_bitwidth: int = 32, transpose: bool = True):
    """Inverse of pack_over_cols; returns uint8 codes of shape (rows, cols*eps)."""
    Wp = np.asarray(W_p)
    if transpose:
        Wp = Wp.T
    eps = packing_bitwidth // nbits
    u = np.ascontiguousarray(Wp).view(_UWORD[packing_bitwidth]).astype(np.uint64)
    shifts = (np.arange(eps, dtype=np.uint64) * np.uint64(nbits))
    codes = (u[:, :, None] >> shifts) & np.uint64((1 << nbits) - 1)
    return codes.reshape(u.shape[0], -1).astype(np.uint8)


# ============================================================================================
# pack-time layout + mode resolution  (core.py:231-299, 336-519)
# ============================================================================================
class Packed:
    """Plain record of what the reference module holds after pack()."""

    def __init__(self):
        self.W_q = None            # torch tensor, (K/eps, N) packed words or (K, N) view
        self.scales = None
        self.zeros = None
        self.bias = None
        self.W_nbits = 0
        self.group_size = 0
        self.unpack_mask = 0
        self.elements_per_sample = 0
        self.input_dtype = FP16
        self.output_dtype = FP16
        self.acc_dtype = FP32
        self.meta_dtype = FP16
        self.channel_scale_mode = 0
        self.W_group_mode = 0
        self.data_contiguous = False
        self.scaled_activations = False
        self.in_features = 0
        self.out_features = 0

    def meta_args(self):
        return [int(self.scaled_activations), self.W_nbits, self.group_size, self.unpack_mask,
                self.elements_per_sample, self.input_dtype, self.output_dtype, self.acc_dtype,
                self.meta_dtype, self.channel_scale_mode, self.W_group_mode, int(self.data_contiguous)]


def default_acc_dtype(input_dtype: int) -> int:
    """core.py:39-54 evaluated on a B200 (fp16 accumulates in fp32 there)."""
    return INT32 if input_dtype == INT8 else FP32


def prepare(W_q, scales, zeros, *, W_nbits, group_size, in_features, out_features,
            input_dtype=FP16, output_dtype=FP16, acc_dtype=None, scaled_activations=False,
            bias=None, fma_mode=True, contiguous=None, packing_bitwidth=32) -> Packed:
    p = Packed()
    if W_nbits not in (1, 2, 4, 8, 16):
        raise NotImplementedError("W_nbits")
    if (in_features % 32 != 0) or (group_size is not None and in_features % group_size != 0):
        raise NotImplementedError("shape")
    if group_size is not None and group_size < 16:
        raise NotImplementedError("group_size")
    group_size = 1 if group_size is None else group_size
    p.W_nbits, p.group_size = W_nbits, group_size
    p.in_features, p.out_features = in_features, out_features
    p.unpack_mask = 2 ** W_nbits - 1
    p.input_dtype, p.output_dtype = input_dtype, output_dtype
    p.acc_dtype = default_acc_dtype(input_dtype) if acc_dtype is None else acc_dtype
    p.meta_dtype = input_dtype
    compute_dtype = TORCH_OF[input_dtype]
    p.scaled_activations = False if compute_dtype in (torch.float16, torch.bfloat16, torch.float32) \
        else bool(scaled_activations)
    is_mx = input_dtype in MX_VALUES

    # core.py:349-357
    if zeros is not None and input_dtype == INT8:
        if isinstance(zeros, torch.Tensor):
            if zeros.float().mean() != zeros.int().float().mean():
                raise Exception("INT8 inputs is not compatible with floating-point zeros.")
        elif isinstance(zeros, float):
            raise Exception("INT8 inputs is not compatible with floating-point zeros.")
    if is_mx:
        packing_bitwidth = 8

    W_q = torch.as_tensor(W_q)
    if W_q.dtype == torch.int8 or W_q.is_floating_point():          # core.py:369-381
        p.W_q = W_q.t()
        p.elements_per_sample = 1
        if contiguous is None:
            contiguous = False
    elif W_q.dtype == torch.uint8:                                    # core.py:384-398
        words, eps = pack_over_cols(W_q.view(out_features, in_features).numpy(), W_nbits,
                                    packing_bitwidth, transpose=True)
        p.W_q = torch.from_numpy(np.ascontiguousarray(words.T)).t()   # (K/eps, N) view of (N, K/eps)
        p.elements_per_sample = eps
        if contiguous is None:
            contiguous = not is_mx
    else:
        raise Exception("Weights were not packed, please check your W_q.dtype")

    p.bias = bias
    p.W_group_mode, p.channel_scale_mode = -1, 0
    if scales is None and zeros is None:                              # core.py:412-416
        p.W_group_mode = 0
        p.channel_scale_mode = 2 if p.scaled_activations else 0
    s = scales.view(out_features, -1).t() if scales is not None else None
    channelwise = False if s is None else (s.numel() == out_features)
    z = None
    if zeros is None:                                                 # core.py:428-445
        p.W_group_mode = 2 if s is not None else 0
    elif isinstance(zeros, torch.Tensor):
        if fma_mode and not channelwise:
            z = (-zeros.float() * scales.float()).to(zeros.dtype).view(out_features, -1).t()
            p.W_group_mode = 4
        else:
            z = zeros.view(out_features, -1).t()
            p.W_group_mode = 3
    else:
        z = int(zeros)
        p.W_group_mode = 3 if s is not None else 1
    assert p.W_group_mode > -1
    if (not p.scaled_activations) and channelwise:                    # core.py:449-464
        p.channel_scale_mode = 1
        p.W_group_mode = 1 if z is not None else 0
    if p.scaled_activations and not channelwise:
        p.channel_scale_mode = 2
    if p.scaled_activations and channelwise:
        p.channel_scale_mode = 3
        p.W_group_mode = 1 if z is not None else 0
    if p.channel_scale_mode in (1, 3):
        assert p.W_group_mode not in (3, 4)

    if isinstance(z, int):                                            # core.py:470-475
        z = torch.tensor(z, dtype=torch.int32)
    if z is None:
        z = torch.tensor([[]], dtype=torch.int32)
    if s is None:
        s = torch.tensor([[]], dtype=torch.int32)
    p.data_contiguous = bool(contiguous)
    if contiguous:
        p.W_q = p.W_q.contiguous()
    s, z = s.contiguous(), z.contiguous()
    if input_dtype in (MXFP16, MXBF16, MXFP8, MXFP4):                 # core.py:489-497
        s = s.to(torch.float8_e8m0fnu).view(torch.uint8)
    if input_dtype == NVFP4:
        s = s.to(torch.float8_e4m3fn)
    if is_mx:
        s = s.T
        p.W_group_mode, p.channel_scale_mode = 2, 0
    p.meta_dtype = DTYPE_OF[s.dtype]
    p.scales, p.zeros = s, z
    return p


# ============================================================================================
# dequantize  (triton_kernels/utils.py:57-89) -> (K, N) tensor in the activation dtype
# ============================================================================================
def _round_to(t64: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """One correctly-rounded conversion of an exact float64 value into `dtype`, back in float64."""
    if dtype == torch.float32:
        return t64.to(torch.float32).to(torch.float64)
    # float64 -> fp16/bf16 directly is a single rounding
    return t64.to(dtype).to(torch.float64)


def unpack_codes(p: Packed) -> torch.Tensor:
    """(K, N) integer codes / raw 8,16-bit weights as float64 (exact)."""
    if p.elements_per_sample > 1:
        words = p.W_q.contiguous().numpy()                            # (K/eps, N)
        bits = words.dtype.itemsize * 8
        u = words.view(_UWORD[bits]).astype(np.uint64)
        eps = p.elements_per_sample
        shifts = (np.arange(eps, dtype=np.uint64) * np.uint64(p.W_nbits))
        codes = (u[:, None, :] >> shifts[None, :, None]) & np.uint64(p.unpack_mask)
        codes = codes.reshape(u.shape[0] * eps, u.shape[1])           # k = word*eps + j
        return torch.from_numpy(codes.astype(np.float64))
    return p.W_q.to(torch.float64)


def dequantize_weights(p: Packed) -> torch.Tensor:
    K, N = p.in_features, p.out_features
    b = unpack_codes(p)                                               # (K, N) float64, exact
    mode = p.W_group_mode
    in_t = TORCH_OF[p.input_dtype]
    meta_t = TORCH_OF[p.meta_dtype]
    meta_is_float = meta_t in (torch.float16, torch.bfloat16, torch.float32)
    gs = p.group_size

    def rows(t):                                                      # (G, N) -> (K, N), g = k // gs
        t = t.to(torch.float64)
        if t.shape[0] == 1:
            return t.expand(K, N)
        return t.repeat_interleave(gs, dim=0)

    zero_is_scalar = p.zeros.numel() == 1
    if mode == 0:
        w = b
    elif mode == 1:                                                   # b.to(meta) - zeros
        zz = p.zeros.to(torch.float64) if zero_is_scalar else rows(p.zeros)
        w = b - zz
        if meta_is_float:
            w = _round_to(w, meta_t)
    elif mode == 2:                                                   # b.to(meta) * scales
        w = _round_to(b * rows(p.scales), meta_t)
    elif mode == 3:
        if zero_is_scalar:                                            # (b - z).to(meta) * scales
            w = _round_to((b - p.zeros.to(torch.float64)) * rows(p.scales), meta_t)
        else:                                                         # (b.to(meta) - z) * scales
            w = _round_to(_round_to(b - rows(p.zeros), meta_t) * rows(p.scales), meta_t)
    elif mode == 4:                                                   # fma(b, s, z'): one rounding
        w = _round_to(b * rows(p.scales) + rows(p.zeros), meta_t)
    else:
        raise ValueError(mode)
    # cast to the dtype the multiply consumes (gemm_kernels.py:384  b.to(input_dtype))
    if in_t == torch.int8:
        return torch.trunc(w).clamp(-128, 127)                        # fp->int8 conversion of integral values
    if in_t in (torch.float8_e4m3fn, torch.float8_e5m2):
        return w.to(torch.float32).to(in_t).to(torch.float64)
    return _round_to(w, in_t)


# ============================================================================================
# MX weights  (gemm_kernels.py:482-531, quant_utils.py:27-37)
# ============================================================================================
def e8m0_to_float(e: np.ndarray) -> np.ndarray:
    return np.exp2(e.astype(np.float64) - 127.0)


def mx_dequantize_weights(p: Packed) -> torch.Tensor:
    """(K, N) float64 exact values of e2m1/e4m3 codes times their 2^(e-127) group scale."""
    K, N, gs = p.in_features, p.out_features, p.group_size
    if p.elements_per_sample == 2:                                    # e2m1 nibbles, even k = low nibble
        codes = unpack_codes(p).numpy().astype(np.int64)
        vals = FP4_VALUES[codes].astype(np.float64)
    else:                                                             # e4m3 bytes
        vals = p.W_q.to(torch.float32).numpy().astype(np.float64)
    s = p.scales                                                      # (N, K/gs) view
    if s.dtype == torch.uint8:
        sc = e8m0_to_float(s.numpy())
    else:                                                             # NVFP4: e4m3 scales
        sc = s.to(torch.float32).numpy().astype(np.float64)
    sc = np.repeat(sc.T, gs, axis=0)                                  # (K, N)
    return torch.from_numpy(vals * sc)


# ============================================================================================
# activation quantisation  (quant_utils.py:231-305)
# ============================================================================================
def quantize_activations_per_token(x: torch.Tensor, w_dtype: torch.dtype):
    """fp32 math; int8 rounds half away from zero (libdevice.round, quant_utils.py:255-257)."""
    xf = x.reshape(-1, x.shape[-1]).to(torch.float32)
    if w_dtype.is_floating_point:
        lo, hi = torch.finfo(w_dtype).min, torch.finfo(w_dtype).max
    else:
        lo, hi = torch.iinfo(w_dtype).min, torch.iinfo(w_dtype).max
    s = xf.abs().amax(dim=1, keepdim=True) / float(hi)
    s = torch.clamp(s, min=1e-6)
    q = torch.clamp(xf / s, lo, hi)
    if not w_dtype.is_floating_point:
        q = torch.sign(q) * torch.floor(q.abs() + 0.5)
    return q.to(w_dtype).view(x.shape), s


# MX / NV activation quantisers: the *_triton_v2 functions forward_functional calls (quant_utils.py:381-391 bitwise
# next-power-of-two, :503-591 mxfp8, :770-857 mxfp4, :860-954 nvfp4).  All arithmetic in fp32, IEEE division (the GPU
# build of the reference divides with div.full.f32, <= 2 ulp -- it can land on the other side of an exact tie).
FP4_THR_POS = np.array([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 7.0], dtype=np.float32)   # quant_utils.py:61-67 (last = pad)


def _next_pow2_exponent(val: np.ndarray, eps_exp: int = -30) -> np.ndarray:
    """quant_utils.py:381-391: biased exponent of the smallest power of two >= val, clamped to [127 + eps_exp, 254]."""
    xi = val.astype(np.float32).view(np.uint32)
    exp = ((xi >> 23) & 0xFF).astype(np.int64)
    exp = exp + ((xi & 0x7FFFFF) != 0)
    return np.clip(exp, 127 + eps_exp, 254)


def _rows_padded(x: torch.Tensor, group: int):
    x2 = x.reshape(-1, x.shape[-1]).to(torch.float32).numpy()
    M, K = x2.shape
    Mp = M + (group - M % group) % group
    xp = np.zeros((Mp, K), np.float32)                                # rows >= M load as 0 (masked load, other=0)
    xp[:M] = x2
    return xp, M, K


def _fp4_codes(wq: np.ndarray) -> np.ndarray:
    """index = #thresholds strictly below |wq| (ties go down), +8 when wq < 0 (so a negative that rounds to 0 is
    code 8 and -0.0 is code 0), then two codes per byte, even k in the low nibble  (quant_utils.py:806-812)."""
    idx = (np.abs(wq)[..., None] > FP4_THR_POS).sum(axis=-1).astype(np.uint8)
    codes = np.where(wq >= 0, idx, idx + 8).astype(np.uint8)
    return (codes[:, 0::2] | (codes[:, 1::2] << 4)).astype(np.uint8)


def quantize_activations_mxfp8(x: torch.Tensor, w_dtype: torch.dtype = torch.float8_e4m3fn):
    """-> (q (M, K) fp8, scales (M padded to 32, K/32) uint8 e8m0 exponents); rows >= M hold 97 (= 2^-30)."""
    xp, M, K = _rows_padded(x, 32)
    g = xp.reshape(xp.shape[0], K // 32, 32)
    hi = np.float32(torch.finfo(w_dtype).max)
    exp = _next_pow2_exponent(np.abs(g).max(axis=2) / hi)
    sc = (exp.astype(np.uint32) << 23).view(np.float32)
    q = np.clip(g / sc[..., None], -hi, hi).reshape(xp.shape)[:M]
    return torch.from_numpy(q).to(w_dtype).view(*x.shape), torch.from_numpy(exp.astype(np.uint8))


def quantize_activations_mxfp4(x: torch.Tensor):
    """-> (codes (M, K/2) uint8, scales (M padded to 32, K/32) uint8 e8m0 exponents)."""
    xp, M, K = _rows_padded(x, 32)
    g = xp.reshape(xp.shape[0], K // 32, 32)
    exp = _next_pow2_exponent(np.abs(g).max(axis=2) / np.float32(6.0))
    sc = (exp.astype(np.uint32) << 23).view(np.float32)
    wq = (g / sc[..., None]).reshape(xp.shape)[:M]
    return torch.from_numpy(_fp4_codes(wq)), torch.from_numpy(exp.astype(np.uint8))


NVFP4_META_SCALE = 0.05                                                  # quant_utils.py:21


def quantize_activations_nvfp4(x: torch.Tensor):
    """-> (codes (M, K/2) uint8, scales (M padded to 16, K/16) float8_e4m3fn); effective scale = e4m3 * 0.05."""
    xp, M, K = _rows_padded(x, 16)
    g = xp.reshape(xp.shape[0], K // 16, 16)
    s = np.abs(g).max(axis=2) / np.float32(6.0 * NVFP4_META_SCALE)
    s8 = torch.from_numpy(np.minimum(s, np.float32(448.0))).to(torch.float8_e4m3fn)
    full = np.maximum(s8.to(torch.float32).numpy() * np.float32(NVFP4_META_SCALE), np.float32(1e-6))
    wq = (g / full[..., None]).reshape(xp.shape)[:M]
    return torch.from_numpy(_fp4_codes(wq)), s8


def mx_dequantize_activations(q: torch.Tensor, scales: torch.Tensor, input_dtype: int) -> torch.Tensor:
    """(M, K) float64 exact values of MX / NV quantised activations (tl.dot_scaled lhs semantics)."""
    if input_dtype == MXFP8:
        vals = q.reshape(-1, q.shape[-1]).to(torch.float32).numpy().astype(np.float64)
    else:
        b = q.reshape(-1, q.shape[-1]).numpy()
        codes = np.stack([b & 15, b >> 4], axis=2).reshape(b.shape[0], -1)
        vals = FP4_VALUES[codes].astype(np.float64)
    M, K = vals.shape
    if scales.dtype == torch.uint8:
        sc = e8m0_to_float(scales.numpy()[:M])
    else:
        sc = scales.to(torch.float32).numpy().astype(np.float64)[:M]
    return torch.from_numpy(vals * np.repeat(sc, K // sc.shape[1], axis=1))


def quantize_activations_for(p: "Packed", x: torch.Tensor):
    """The quantiser forward_functional picks from (input_dtype, channel_scale_mode)  (core.py:157-175)."""
    if p.input_dtype in (INT8, FP8, FP8E5):
        return quantize_activations_per_token(x, TORCH_OF[p.input_dtype])
    if p.input_dtype == MXFP8 and p.channel_scale_mode == 4:
        return quantize_activations_mxfp8(x)
    if p.input_dtype == MXFP8 and p.channel_scale_mode == 2:
        return quantize_activations_per_token(x, torch.float8_e4m3fn)
    if p.input_dtype == MXFP4 and p.channel_scale_mode == 4:
        return quantize_activations_mxfp4(x)
    if p.input_dtype == NVFP4 and p.channel_scale_mode == 4:
        return quantize_activations_nvfp4(x)
    return x, None                                                    # no branch taken: x goes in as it is


# ============================================================================================
# forward  (core.py:128-195 + gemm_kernels.py:347-413)
# ============================================================================================
def forward(p: Packed, x: torch.Tensor, scales_x: torch.Tensor | None = None) -> torch.Tensor:
    """x: (..., K).  Returns (..., N) in the output dtype.  Dynamic-activation layers quantise x
    here unless `scales_x` is given (then x is taken as already quantised)."""
    out_t = TORCH_OF[p.output_dtype]
    lead = x.shape[:-1]
    in_t = TORCH_OF[p.input_dtype]
    if p.scaled_activations and scales_x is None:
        x, scales_x = quantize_activations_for(p, x)
        lead = lead if x.dim() == len(lead) + 1 else x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if p.input_dtype in MX_VALUES:
        # gemm_kernels.py:482-547: acc (fp32) += dot_scaled(a, scales_a, b, scales_b); NVFP4: acc *= 0.05^2;
        # channel_scale_mode 2: acc.to(out dtype) * scales_a (fp32), stored in the out dtype
        W = mx_dequantize_weights(p)
        if p.scaled_activations and p.channel_scale_mode == 4:
            xv = mx_dequantize_activations(x2, scales_x, p.input_dtype)
        else:
            xv = x2.to(torch.float64) if x2.dtype != torch.float8_e4m3fn else x2.to(torch.float32).to(torch.float64)
        acc = (xv @ W).to(torch.float32)
        if p.group_size == 16:
            acc = acc * (0.05 ** 2)
        if p.scaled_activations and p.channel_scale_mode == 2:
            sa = scales_x.reshape(-1, 1).to(torch.float32)
            y = (acc.to(out_t).to(torch.float32) * sa).to(out_t)
        else:
            y = acc.to(out_t)
    else:
        W = dequantize_weights(p)                                     # (K, N) float64
        acc = x2.to(torch.float64) @ W                                # exact sum, then one rounding
        if p.acc_dtype == INT32:
            acc_v = acc                                               # integers, exact
        else:
            acc_v = acc.to(torch.float32).to(torch.float64)
        meta_t = TORCH_OF[p.meta_dtype]
        csm = p.channel_scale_mode
        if csm == 0:
            y = acc_v.to(out_t) if out_t.is_floating_point else acc_v.to(out_t)
        else:
            a_m = _round_to(acc_v, meta_t) if meta_t.is_floating_point else acc_v
            if csm == 1:
                sb = p.scales.reshape(1, -1).to(torch.float64)
                r = a_m * sb
                r = _round_to(r, meta_t)
            else:
                sa = scales_x.reshape(-1, 1).to(torch.float64)
                if csm == 2:
                    sc = _round_to(sa * 1.0, _promote(scales_x.dtype, meta_t))
                else:
                    sb = p.scales.reshape(1, -1).to(torch.float64)
                    sc = _round_to(sa * sb, _promote(scales_x.dtype, meta_t))
                r = _round_to(a_m * sc, _promote(scales_x.dtype, meta_t))
            y = r.to(out_t)
    y = y.view(*lead, p.out_features)
    if p.bias is not None:
        y = (y.to(torch.float64) + p.bias.to(torch.float64)).to(out_t)   # out += bias  (core.py:192-193)
    return y


def _promote(a: torch.dtype, b: torch.dtype) -> torch.dtype:
    if not b.is_floating_point:
        return a
    return torch.promote_types(a, b)


# ============================================================================================
# MXFP weight quantiser, window_size = 0  (quant_utils.py:96-163)
# ============================================================================================
def quantize_mxfp4(W: torch.Tensor):
    """returns (codes uint8 (N*K/32, 32) e2m1 indices, scales float8_e8m0fnu (N*K/32, 1))."""
    Wf = W.reshape(-1, 32).to(torch.float32)
    ideal = Wf.abs().amax(dim=1, keepdim=True) / 6.0
    sc = torch.pow(2.0, torch.ceil(torch.log2(ideal))).clamp_(min=2.0 ** -30)
    v = (Wf / sc)
    idx = torch.bucketize(v.abs(), torch.from_numpy(FP4_THRESHOLDS), right=False)   # searchsorted(left)
    mag = torch.from_numpy(FP4_VALUES[:8])[idx]
    q = mag * torch.sign(v)
    # to_index(): first table entry equal to the value (so -0.0 == 0 -> code 0)
    table = torch.from_numpy(FP4_VALUES)
    codes = (q.reshape(-1, 1) == table.reshape(1, -1)).to(torch.uint8).argmax(dim=1).to(torch.uint8)
    return codes.view(Wf.shape), sc.to(torch.float8_e8m0fnu)


def quantize_mxfp8(W: torch.Tensor, fp8=torch.float8_e4m3fn):
    Wf = W.reshape(-1, 32).to(torch.float32)
    mx = torch.finfo(fp8).max
    ideal = Wf.abs().amax(dim=1, keepdim=True) / mx
    sc = torch.pow(2.0, torch.ceil(torch.log2(ideal))).clamp_(min=2.0 ** -30)
    q = (Wf / sc).clamp_(-mx, mx).to(fp8)
    return q, sc.to(torch.float8_e8m0fnu)


def quantize_nvfp4(W: torch.Tensor):
    """quant_utils.py:165-211, window_size = 0: group 16, e4m3 scales, effective scale = e4m3 * 0.05 (>= 1e-6).
    returns (codes uint8 (N*K/16, 16), scales float8_e4m3fn (N*K/16, 1))."""
    Wf = W.reshape(-1, 16).to(torch.float32)
    ideal = Wf.abs().amax(dim=1, keepdim=True)
    ideal = ideal / 6
    ideal = ideal / NVFP4_META_SCALE
    sc = ideal.clamp(max=448.0).to(torch.float8_e4m3fn)
    full = (sc.to(torch.float32) * NVFP4_META_SCALE).clamp(min=1e-6)
    v = Wf / full
    idx = torch.bucketize(v.abs(), torch.from_numpy(FP4_THRESHOLDS), right=False)
    q = torch.from_numpy(FP4_VALUES[:8])[idx] * torch.sign(v)
    table = torch.from_numpy(FP4_VALUES)
    codes = (q.reshape(-1, 1) == table.reshape(1, -1)).to(torch.uint8).argmax(dim=1).to(torch.uint8)
    return codes.view(Wf.shape), sc


# ============================================================================================
# CPU baseline used by bench.py: torch dequant + matmul on host cores
# ============================================================================================
def cpu_dequant_matmul_fast(Wp_i32: torch.Tensor, scales: torch.Tensor, zeros_fma: torch.Tensor,
                            x: torch.Tensor, group_size: int) -> torch.Tensor:
    """A16W4 (mode 4) on host cores without materialising the dequantised matrix: per group g
        y[m, n] = sum_g s[g, n] * (sum_{k in g} x[m, k] q[k, n])  +  z'[g, n] * sum_{k in g} x[m, k]
    with the codes taken straight from the packed bytes (a word (r, n) holds k = 8r .. 8r+7, byte b = nibbles 2b, 2b+1;
    bitpack.py:42-54).  fp32 throughout; same result as cpu_dequant_matmul up to fp32 summation order.  This is the
    timed CPU baseline of bench.py (the plain version below builds a (K/8, 8, N) int32 shift tensor per call)."""
    Kp, N = Wp_i32.shape
    G = scales.shape[0]
    rpg = Kp // G                                                  # packed rows per group
    M = x.shape[0]
    by = Wp_i32.contiguous().view(torch.uint8).view(G, rpg, N, 4)  # little endian: byte b of the word
    xf = x.to(torch.float32).view(M, G, rpg, 4, 2)                 # [m, g, r, b, h] = x[m, 8 (g rpg + r) + 2 b + h]
    lo = (by & 15).to(torch.float32)
    hi = (by >> 4).to(torch.float32)
    P = torch.einsum("mgrb,grnb->mgn", xf[..., 0], lo) + torch.einsum("mgrb,grnb->mgn", xf[..., 1], hi)
    sx = xf.sum(dim=(2, 3, 4))                                     # (M, G)
    return (P * scales.to(torch.float32).unsqueeze(0)).sum(1) + sx @ zeros_fma.to(torch.float32)


def cpu_dequant_matmul(Wp_i32: torch.Tensor, scales: torch.Tensor, zeros_fma: torch.Tensor,
                       x: torch.Tensor, group_size: int) -> torch.Tensor:
    """A16W4 (mode 4) dequant + matmul in fp32 with plain torch ops, vectorised (the timed baseline)."""
    Kp, N = Wp_i32.shape
    shifts = torch.arange(0, 32, 4, dtype=torch.int32).view(1, 8, 1)
    codes = ((Wp_i32.view(Kp, 1, N) >> shifts) & 15).reshape(Kp * 8, N).to(torch.float32)
    s = scales.to(torch.float32).repeat_interleave(group_size, dim=0)
    z = zeros_fma.to(torch.float32).repeat_interleave(group_size, dim=0)
    W = codes * s + z
    return x.to(torch.float32) @ W
