"""Activation quantisation ops (CUDA, through the C ABI) and the MXFP/NVFP weight quantiser.

Reference counterparts: gemlite/quant_utils.py:308-347 (`scale_activations_per_token`, a Triton
kernel there; `glb_quant_act_per_token` here) and :70-225 (`WeightQuantizerMXFP`, load-time torch
code in both; here written without per-device lookup tables so it runs on whatever device holds the
weights)."""
from typing import Tuple

import torch
from torch import Tensor

from . import _cabi

NVFP4_META_SCALE = 0.05


def get_dtype_range(compute_dtype: torch.dtype):
    info = torch.finfo(compute_dtype) if compute_dtype.is_floating_point else torch.iinfo(compute_dtype)
    return info.min, info.max


def scale_activations_per_token(tensor: Tensor, w_dtype: torch.dtype, fp32_scale: bool = True) -> Tuple[Tensor, Tensor]:
    """Per-row dynamic quantisation to int8 / fp8: returns (q with tensor's shape, scales (M, 1) fp32)."""
    _cabi.require_cuda(tensor, "scale_activations_per_token")
    if not fp32_scale:
        raise NotImplementedError("only fp32 activation scales are supported")
    x_shape = tensor.shape
    x2 = tensor.reshape(-1, x_shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, K = x2.shape
    q = torch.empty((M, K), dtype=w_dtype, device=tensor.device)
    s = torch.empty((M, 1), dtype=torch.float32, device=tensor.device)
    if M > 0:
        with torch.cuda.device(tensor.device):
            rc = _cabi.load().glb_quant_act_per_token(x2.data_ptr(), M, K, _cabi.dtype_tag(x2.dtype),
                                                      _cabi.dtype_tag(w_dtype), q.data_ptr(), s.data_ptr(),
                                                      _cabi.current_stream_ptr(tensor.device))
        _cabi.check(rc, "glb_quant_act_per_token")
    return q.view(x_shape), s


scale_activations_per_token_cuda = scale_activations_per_token


def _quant_act_mx(tensor: Tensor, mx_tag: int, q_dtype: torch.dtype, group: int, scale_dtype: torch.dtype, what: str):
    """x (..., K) -> (q (M, K or K/2), scales (M padded to `group` rows, K/group)) through glb_quant_act_mx.
    Like the reference's *_triton_v2 functions the result is 2-D (quant_utils.py:546-591, 819-857, 917-954)."""
    _cabi.require_cuda(tensor, what)
    x2 = tensor.reshape(-1, tensor.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, K = x2.shape
    M_pad = (M + group - 1) // group * group
    q = torch.empty((M, K if q_dtype != torch.uint8 else K // 2), dtype=q_dtype, device=tensor.device)
    scales = torch.empty((M_pad, K // group), dtype=scale_dtype, device=tensor.device)
    if M > 0:
        with torch.cuda.device(tensor.device):
            rc = _cabi.load().glb_quant_act_mx(x2.data_ptr(), M, K, _cabi.dtype_tag(x2.dtype), mx_tag, q.data_ptr(),
                                               scales.data_ptr(), _cabi.current_stream_ptr(tensor.device))
        _cabi.check(rc, "glb_quant_act_mx")
    return q, scales


def scale_activations_mxfp8(tensor: Tensor, w_dtype: torch.dtype = torch.float8_e4m3fn) -> Tuple[Tensor, Tensor]:
    """(M, K) e4m3 + (M padded to 32, K/32) uint8 e8m0 exponents."""
    if w_dtype != torch.float8_e4m3fn:
        raise NotImplementedError("scale_activations_mxfp8: only float8_e4m3fn activations are supported")
    return _quant_act_mx(tensor, 16, torch.float8_e4m3fn, 32, torch.uint8, "scale_activations_mxfp8")      # DType.MXFP8


def scale_activations_mxfp4(tensor: Tensor) -> Tuple[Tensor, Tensor]:
    """(M, K/2) e2m1 codes (even k = low nibble) + (M padded to 32, K/32) uint8 e8m0 exponents."""
    return _quant_act_mx(tensor, 17, torch.uint8, 32, torch.uint8, "scale_activations_mxfp4")              # DType.MXFP4


def scale_activations_nvfp4(tensor: Tensor) -> Tuple[Tensor, Tensor]:
    """(M, K/2) e2m1 codes + (M padded to 16, K/16) float8_e4m3fn scales (effective scale = e4m3 * 0.05)."""
    return _quant_act_mx(tensor, 18, torch.uint8, 16, torch.float8_e4m3fn, "scale_activations_nvfp4")      # DType.NVFP4


# --------------------------------------------------------------------------------------------------
# weight quantiser (load time)
# --------------------------------------------------------------------------------------------------
_FP4_POS = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
_FP4_THR = [0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0]    # midpoints; ties go to the lower magnitude
_FP4_ALL = _FP4_POS + [-v for v in _FP4_POS]            # index == e2m1 bit pattern (sign = bit 3)


class WeightQuantizerMXFP:
    """MXFP8 / MXFP4 (group 32, e8m0 scales) and NVFP4 (group 16, e4m3 scales) weight quantisers."""

    def __init__(self, compute_dtype=torch.bfloat16, device="cuda:0"):
        self.compute_dtype = compute_dtype
        self.device = device

    @staticmethod
    def round_to_closest_fp4(t: Tensor) -> Tensor:
        thr = torch.tensor(_FP4_THR, dtype=t.dtype, device=t.device)
        pos = torch.tensor(_FP4_POS, dtype=t.dtype, device=t.device)
        return pos[torch.bucketize(t.abs(), thr, right=False)] * t.sign()

    @staticmethod
    def to_index(W_q: Tensor) -> Tensor:
        """fp4 values -> e2m1 codes (the first table entry equal to the value, so -0 maps to 0)."""
        assert W_q.is_floating_point(), "Input should be floating point fp4 values."
        mag = W_q.abs()
        thr = torch.tensor(_FP4_THR, dtype=mag.dtype, device=mag.device)
        idx = torch.bucketize(mag, thr, right=False)
        neg = (W_q < 0) & (mag > 0)
        return (idx + 8 * neg.to(idx.dtype)).to(torch.uint8)

    def quantize_mxfp8(self, W: Tensor, index: bool = False, mx_fp8_dtype: torch.dtype = torch.float8_e4m3fn):
        lo, hi = get_dtype_range(mx_fp8_dtype)
        Wf = W.reshape(-1, 32).float()
        ideal = Wf.abs().amax(dim=1, keepdim=True) / hi
        scales = torch.pow(2.0, torch.ceil(torch.log2(ideal))).clamp_(min=2.0 ** -30)
        W_q = (Wf / scales).clamp_(min=lo, max=hi).to(mx_fp8_dtype)
        if not index:
            W_q = W_q.to(Wf.dtype)
        return W_q, scales.to(torch.float8_e8m0fnu)

    def quantize_mxfp4(self, W: Tensor, window_size: int = 0, index: bool = False):
        eps = 2.0 ** -30
        Wf = W.reshape(-1, 32).float()
        ideal = Wf.abs().amax(dim=1, keepdim=True) / 6.0
        log2s = torch.ceil(torch.log2(ideal))
        if window_size == 0:
            scales = torch.pow(2.0, log2s)
        else:   # pick the power-of-two scale in a window that minimises the mean abs error
            offs = torch.arange(-window_size, window_size + 1, device=W.device, dtype=log2s.dtype).view(1, -1)
            cand = torch.pow(2.0, log2s + offs).clamp_(min=eps)
            q = self.round_to_closest_fp4(Wf.unsqueeze(1) / cand.unsqueeze(-1))
            err = (Wf.unsqueeze(1) - q * cand.unsqueeze(-1)).abs().mean(dim=-1)
            scales = torch.gather(cand, 1, torch.argmin(err, dim=1, keepdim=True))
        scales = scales.clamp_(eps)
        W_q = self.round_to_closest_fp4(Wf / scales)
        if index:
            W_q = self.to_index(W_q)
        return W_q, scales.to(torch.float8_e8m0fnu)

    def quantize_nvfp4(self, W: Tensor, window_size: int = 0, index: bool = False):
        fp8 = torch.float8_e4m3fn
        eps = 1e-6
        Wf = W.reshape(-1, 16).float()
        ideal = Wf.abs().amax(dim=1, keepdim=True) / 6.0 / NVFP4_META_SCALE
        scales = ideal.clamp_(max=torch.finfo(fp8).max).to(fp8)
        if window_size != 0:
            # neighbouring e4m3 bit patterns of the ideal scale (quant_utils.py:186-205); the error is measured against
            # candidate * q WITHOUT the 0.05 meta scale, as the reference does
            offs = torch.arange(-window_size, window_size + 1, device=W.device, dtype=torch.int).view(1, -1)
            cand = (scales.view(torch.int8) + offs).clamp_(-128, 127).to(torch.int8)
            cand[cand == -1] = 1            # the two e4m3 NaN patterns
            cand[cand == 127] = 1
            cand = cand.view(fp8).float()
            cand[cand < eps] = eps
            q = self.round_to_closest_fp4(Wf.unsqueeze(1) / (cand * NVFP4_META_SCALE).unsqueeze(-1))
            err = (Wf.unsqueeze(1) - q * cand.unsqueeze(-1)).abs().mean(dim=-1)
            scales = torch.gather(cand, 1, torch.argmin(err, dim=1, keepdim=True)).to(fp8)
        full = (scales.to(Wf.dtype) * NVFP4_META_SCALE).clamp_(min=eps)
        W_q = self.round_to_closest_fp4(Wf / full)
        if index:
            W_q = self.to_index(W_q)
        return W_q, scales

    def dequantize(self, W_q, scales, shape=None, dtype=None):
        if W_q.dtype == torch.uint8:
            W_q = torch.tensor(_FP4_ALL, dtype=torch.float32, device=W_q.device)[W_q.int()]
        group = W_q.numel() // scales.numel()
        out = W_q.reshape(-1, group).float() * scales.float().reshape(-1, 1)
        if shape is not None:
            out = out.view(shape)
        return out.to(self.compute_dtype if dtype is None else dtype)
