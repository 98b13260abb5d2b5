"""GemLiteLinear: the drop-in module (pack / forward / state_dict) on top of the sm_100a C ABI.

Interface parity with the reference module (gemlite/core.py): constructor arguments and validation
(:231-299), `pack()` layout + mode resolution (:336-519), `get_tensor_args()/get_meta_args()`
(:522-538), `forward` / `forward_manual` (:541-557), `state_dict` keys and `load_state_dict`
(:301-333, :502-517), the `gemlite::forward_functional` custom op with its fake (:128-206) and the
config helpers (:86-97, :559-627).  The bodies are new: packing, activation quantisation and every
matmul run in hand-written CUDA (gemlite_b200/csrc) reached through `_cabi`; there is no Triton,
no autotuner and no CPU path.
"""
from __future__ import annotations

import ctypes
import json
import logging
import os
import threading
from typing import List, Optional, Union

import torch
from torch import Tensor

from . import _cabi
from .bitpack import pack_weights_over_cols
from .dtypes import (DType, DTYPE_TO_TORCH, TORCH_TO_DTYPE, FP8_INT8_DTYPES, is_mx_dtype)
from .quant_utils import (scale_activations_per_token, scale_activations_mxfp8, scale_activations_mxfp4,
                          scale_activations_nvfp4)

logger = logging.getLogger(__name__)
_FILE_LOCK = threading.Lock()

# ---------------------------------------------------------------------------------------------------
# module-level configuration (reference: core.py:39-97, triton_kernels/config.py)
# ---------------------------------------------------------------------------------------------------
# Accumulator type per activation type.  On B200 the tensor cores and the decode kernel accumulate
# fp16/bf16 products in fp32 (the reference also resolves FP16 -> FP32 on this GPU, core.py:40).
GEMLITE_ACC_DTYPE = {
    DType.FP16: DType.FP32, DType.BF16: DType.FP32, DType.FP32: DType.FP32,
    DType.FP8: DType.FP32, DType.FP8e5: DType.FP32, DType.FP8e4nuz: DType.FP32, DType.FP8e5nuz: DType.FP32,
    DType.INT8: DType.INT32,
    DType.MXFP16: DType.FP32, DType.MXBF16: DType.FP32, DType.MXFP8: DType.FP32, DType.MXFP4: DType.FP32,
    DType.NVFP4: DType.FP32,
}

# kernel names of the reference, in its order (core.py:56-66); forward_manual() accepts these
GEMLITE_MATMUL_TYPES = ["GEMV", "GEMV_REVSPLITK", "GEMV_SPLITK", "GEMM_SPLITK", "GEMM"]
GEMLITE_MATMUL_TYPES_MAPPING = {name: i for i, name in enumerate(GEMLITE_MATMUL_TYPES)}
GEMLITE_TRITON_CONFIG_CACHE = {}   # kept as an opaque dict: there is no autotuner to feed on B200

_SETTINGS = {"autotune": "fast", "kernel_caching": False, "m_bucket_fn": None}


def set_autotune(config=None, **kwargs):
    """Accepted for API compatibility; the CUDA kernels have fixed, shape-derived launch parameters."""
    _SETTINGS["autotune"] = config


def set_kernel_caching(enable: bool):
    """Accepted for API compatibility (the reference's output ring buffer has no equivalent here)."""
    _SETTINGS["kernel_caching"] = bool(enable)


def set_autotune_setting(fct):
    _SETTINGS["m_bucket_fn"] = fct


def set_packing_bitwidth(packing_bitwidth: int):
    assert packing_bitwidth in (8, 16, 32, 64), "Unsuported bitpacking width"
    GemLiteLinear.PACKING_BITWIDTH = packing_bitwidth


def set_acc_dtype(dtype):
    assert dtype in [DType.FP16, DType.FP32], "Invalid dtype (should be DType.FP16 or DType.FP32)."
    # fp16 accumulation is never faster than fp32 on sm_100a tensor cores; the request is recorded so
    # that metadata round-trips, the kernels keep accumulating in fp32.
    GEMLITE_ACC_DTYPE[DType.FP16] = dtype


def get_default_gemv(W_nbits: int, mx_dtype: bool = False) -> str:
    if mx_dtype:
        return "GEMM_SPLITK"
    return "GEMV_REVSPLITK" if (W_nbits < 8) else "GEMV_SPLITK"


def get_matmul_type(batch_size: int, W_nbits: int, mx_dtype: bool = False) -> str:
    """Batch-size dispatch of the reference (core.py:108-114)."""
    if batch_size > 64:
        return "GEMM"
    if batch_size > 1:
        return "GEMM_SPLITK"
    return get_default_gemv(W_nbits, mx_dtype)


# ---------------------------------------------------------------------------------------------------
# descriptor plumbing
# ---------------------------------------------------------------------------------------------------
_DESC_CACHE = {}
_DESC_CACHE_MAX = int(os.environ.get("GEMLITE_B200_DESC_CACHE", "65536"))
_WORKSPACE = {}
_WORKSPACE_RETIRED = []
_CAPTURE_WS = {}      # (device, stream) -> (capture id, buffer zeroed inside that capture)
_WS_BYTES = {}       # (descriptor id, M, matmul_type) -> workspace bytes the C library asked for (descriptors are cached, ids are stable)


def _strides2(t: Tensor):
    if t.dim() == 2:
        return t.stride(0), t.stride(1)
    return 0, 0


def _make_desc(W_q: Tensor, scales: Tensor, zeros: Tensor, meta: List[int]) -> _cabi.GlbLinear:
    # pointers alone are not an identity: the allocator hands freed addresses to later layers of other shapes
    key = (W_q.data_ptr(), scales.data_ptr(), zeros.data_ptr(), W_q.shape, W_q.stride(), W_q.dtype,
           scales.shape, scales.stride(), scales.dtype, zeros.shape, zeros.stride(), zeros.dtype, tuple(meta))
    d = _DESC_CACHE.get(key)
    if d is not None:
        if len(_DESC_CACHE) >= _DESC_CACHE_MAX // 2:     # keep recently used entries at the young end
            _DESC_CACHE[key] = _DESC_CACHE.pop(key)
        return d
    d = _cabi.GlbLinear()
    eps = int(meta[4])
    d.W_q = W_q.data_ptr()
    d.wq_rows, d.wq_cols = W_q.shape
    d.wq_stride0, d.wq_stride1 = W_q.stride()
    d.wq_dtype = _cabi.dtype_tag(W_q.dtype)
    d.scales = scales.data_ptr() if scales.numel() else None
    d.scales_numel = scales.numel()
    d.s_stride0, d.s_stride1 = _strides2(scales)
    d.scales_dtype = _cabi.dtype_tag(scales.dtype)
    d.zeros = zeros.data_ptr() if zeros.numel() else None
    d.zeros_numel = zeros.numel()
    d.z_stride0, d.z_stride1 = _strides2(zeros)
    d.zeros_dtype = _cabi.dtype_tag(zeros.dtype)
    d.K = W_q.shape[0] * eps
    d.N = W_q.shape[1]
    (d.scaled_activations, d.W_nbits, d.group_size, d.unpack_mask, d.elements_per_sample, d.input_dtype,
     d.output_dtype, d.acc_dtype, d.meta_dtype, d.channel_scale_mode, d.W_group_mode, d.data_contiguous) = \
        [int(v) for v in meta]
    if len(_DESC_CACHE) >= _DESC_CACHE_MAX:
        # evict the oldest quarter (dicts keep insertion order; a hit re-inserts, see above) instead of dropping the
        # whole cache: a model with more layers than the bound would otherwise rebuild every descriptor every step
        for old in list(_DESC_CACHE)[: _DESC_CACHE_MAX // 4]:
            dd = _DESC_CACHE.pop(old)
            for wk in [k for k in _WS_BYTES if k[0] == id(dd)]:
                _WS_BYTES.pop(wk, None)
    _DESC_CACHE[key] = d
    return d


def _workspace(device, nbytes: int) -> Optional[Tensor]:
    """Per (device, stream) scratch for the split-K launches.  Zero-initialised once (the kernels keep arrival counters
    in it and leave them at zero).  Sized so that it never has to grow (the largest request is < 10 MB: at most one
    work item per SM pair); should it grow anyway, the old buffer stays alive because captured CUDA graphs may still
    point at it.
    Under CUDA-graph capture a buffer allocated here is zeroed by a memset NODE of that graph, i.e. only when that graph
    replays: it is therefore used by that capture alone (keyed by the capture id) -- a second graph captured on the same
    stream, or eager calls, never see it, whichever of them runs first."""
    if nbytes == 0:
        return None
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device, stream)
    ws = _WORKSPACE.get(key)
    if ws is not None and ws.numel() >= nbytes:
        return ws                    # zeroed eagerly (warm-up on this stream): valid inside and outside any capture
    if torch.cuda.is_current_stream_capturing():
        cap = int(_cabi.load().glb_stream_capture_id(stream))
        if cap:
            held = _CAPTURE_WS.get(key)
            if held is None or held[0] != cap or held[1].numel() < nbytes:
                held = (cap, torch.zeros(max(nbytes, 4 << 20), dtype=torch.uint8, device=device))
                _CAPTURE_WS[key] = held                       # the graph's pool keeps earlier captures' buffers alive
            return held[1]
    if ws is not None:
        _WORKSPACE_RETIRED.append(ws)             # captured graphs may still point at it
        del _WORKSPACE_RETIRED[:-8]                # bounded: a buffer is only ever replaced by a larger one (<= 8 doublings)
    ws = torch.zeros(max(nbytes, 16 << 20), dtype=torch.uint8, device=device)
    _WORKSPACE[key] = ws
    return ws


def _matmul(x2d: Tensor, bias: Optional[Tensor], tensor_args: List[Tensor], meta_args: List[int], scales_x,
            matmul_type: int) -> Tensor:
    W_q, scales, zeros = tensor_args
    _cabi.require_cuda(x2d, "forward")
    if W_q.device != x2d.device:
        raise RuntimeError(f"forward: activations are on {x2d.device}, the layer's weights on {W_q.device}")
    lib = _cabi.load()
    desc = _make_desc(W_q, scales, zeros, meta_args)
    M = x2d.shape[0]
    out = torch.empty((M, desc.N), dtype=DTYPE_TO_TORCH[int(meta_args[6])], device=x2d.device)
    if M == 0:
        return out
    if bias is not None and (bias.dtype != out.dtype or not bias.is_contiguous()):
        bias = bias.to(out.dtype).contiguous()
    dev = x2d.device
    switch = dev.index is not None and dev.index != torch.cuda.current_device()
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        wkey = (id(desc), M, matmul_type)
        need = _WS_BYTES.get(wkey)
        if need is None:
            need = lib.glb_workspace_bytes(ctypes.byref(desc), M, matmul_type)
            if len(_WS_BYTES) > 65536:
                _WS_BYTES.clear()
            _WS_BYTES[wkey] = need
        ws = _workspace(dev, need)
        rc = lib.glb_matmul(ctypes.byref(desc), x2d.data_ptr(),
                            scales_x.data_ptr() if scales_x is not None else None, out.data_ptr(), M, matmul_type,
                            ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                            bias.data_ptr() if bias is not None else None, _cabi.current_stream_ptr(dev))
    finally:
        if switch:
            torch.cuda.set_device(prev)
    _cabi.check(rc, "glb_matmul")
    return out


def matmul_bcast(x2d: Tensor, bias: Optional[Tensor], tensor_args: List[Tensor], meta_args: List[int], dst_ptrs, col0: int) -> None:
    """Decode call (M == 1) whose result row is stored at column `col0` of every destination in `dst_ptrs` (a ctypes
    array of device-accessible pointers: this rank's gathered-output row and the peers' symmetric buffers) -- the fused
    output all-gather of `sharding.SymmetricOutputs`.  Raises NotImplementedError when the layer / batch size has no
    multi-destination kernel (the caller then uses forward + a collective)."""
    W_q, scales, zeros = tensor_args
    _cabi.require_cuda(x2d, "matmul_bcast")
    if x2d.shape[0] != 1 or not x2d.is_contiguous():
        raise NotImplementedError("matmul_bcast: one contiguous activation row")
    desc = _make_desc(W_q, scales, zeros, meta_args)
    if x2d.shape[1] != desc.K:
        raise RuntimeError(f"matmul_bcast: activations have {x2d.shape[1]} input features, the layer expects {desc.K}")
    if bias is not None and (bias.dtype != x2d.dtype or not bias.is_contiguous()):
        bias = bias.to(x2d.dtype).contiguous()
    with torch.cuda.device(x2d.device):
        rc = _cabi.load().glb_matmul_bcast(ctypes.byref(desc), x2d.data_ptr(), None, dst_ptrs, len(dst_ptrs), int(col0), 1,
                                           bias.data_ptr() if bias is not None else None, _cabi.current_stream_ptr(x2d.device))
    _cabi.check(rc, "glb_matmul_bcast")


# ---------------------------------------------------------------------------------------------------
# functional forward (custom op so that torch.compile / vLLM can trace through it)
# ---------------------------------------------------------------------------------------------------
_OP_NS = os.environ.get("GEMLITE_B200_OP_NAMESPACE", "gemlite")


def enable_activation_scaling(batch_size):
    return True


def _forward_eager(x: Tensor, bias: Optional[Tensor], tensor_args, meta_args, matmul_type: int = -1) -> Tensor:
    """Body of the functional forward.  `GemLiteLinear.forward` calls it directly in eager mode: the custom-op
    dispatcher below costs ~20 us per call on the host, ten launches' worth at decode sizes; it is only needed when
    a compiler traces the module."""
    out_features = tensor_args[0].shape[1]
    if not x.is_contiguous():
        x = x.contiguous()
    batch_size = x.numel() // x.shape[-1]
    out_shape = x.shape[:-1] + (out_features,)
    input_dtype = DType(int(meta_args[5]))

    scales_x = None
    if bool(meta_args[0]) and enable_activation_scaling(batch_size):
        csm = int(meta_args[9])
        if input_dtype in FP8_INT8_DTYPES:
            x, scales_x = scale_activations_per_token(x, w_dtype=DTYPE_TO_TORCH[input_dtype.value])
        elif input_dtype == DType.MXFP8 and csm == 2:
            x, scales_x = scale_activations_per_token(x, w_dtype=torch.float8_e4m3fn)
        elif input_dtype == DType.MXFP8 and csm == 4:      # core.py:165-175: MX micro-scales for the activations
            x, scales_x = scale_activations_mxfp8(x, w_dtype=torch.float8_e4m3fn)
        elif input_dtype == DType.MXFP4 and csm == 4:
            x, scales_x = scale_activations_mxfp4(x)
        elif input_dtype == DType.NVFP4 and csm == 4:
            x, scales_x = scale_activations_nvfp4(x)
        else:
            raise NotImplementedError(
                f"dynamic activation quantisation for {input_dtype} (channel_scale_mode={csm}) is not implemented")
    x2d = x.view(-1, x.shape[-1])
    expected = DTYPE_TO_TORCH[input_dtype.value]
    if x2d.dtype != expected:
        raise TypeError(f"activations are {x2d.dtype}, layer expects {expected}")
    # the kernels (and their TMA maps) take the row length of x from the layer: a narrower / wider activation would be
    # read out of bounds, so refuse it here.  Packed fp4 activations carry two codes per byte.
    in_features = tensor_args[0].shape[0] * int(meta_args[4])
    k_have = x2d.shape[1] * (2 if input_dtype in (DType.MXFP4, DType.NVFP4) else 1)
    if k_have != in_features:
        raise RuntimeError(f"forward: activations have {k_have} input features, the layer expects {in_features}")
    if x2d.data_ptr() % 16 != 0:       # a contiguous slice of a flat buffer: the TMA / 128-bit paths need 16-byte rows
        x2d = x2d.clone()
    out = _matmul(x2d, bias, tensor_args, meta_args, scales_x, int(matmul_type))
    return out.view(out_shape)


@torch.library.custom_op(f"{_OP_NS}::forward_functional", mutates_args=())
def forward_functional(x: Tensor, bias: Optional[Tensor], tensor_args: List[Tensor], meta_args: List[int],
                       matmul_type: int = -1) -> Tensor:
    """x (..., K) -> (..., N).  matmul_type: -1 auto (reference dispatch), else index into
    GEMLITE_MATMUL_TYPES.  Dynamic-activation layers quantise x per token first (core.py:155-175)."""
    return _forward_eager(x, bias, tensor_args, meta_args, matmul_type)


@torch.library.register_fake(f"{_OP_NS}::forward_functional")
def _forward_functional_fake(x, bias, tensor_args, meta_args, matmul_type=-1):
    out_features = tensor_args[0].shape[1]
    return torch.empty(x.shape[:-1] + (out_features,), device=x.device, dtype=DTYPE_TO_TORCH[int(meta_args[6])])


# ---------------------------------------------------------------------------------------------------
# the module
# ---------------------------------------------------------------------------------------------------
class GemLiteLinear(torch.nn.Module):
    SUPPORTED_BITS_TRITON = [1, 2, 4, 8, 16]
    SUPPORTED_DTYPES = [
        DType.FP16, DType.BF16, DType.FP32, DType.FP8, DType.FP8e4, DType.FP8e4nuz, DType.FP8e5, DType.FP8e5nuz,
        DType.INT8, DType.MXFP16, DType.MXBF16, DType.MXFP8, DType.MXFP4, DType.NVFP4,
    ]
    MIN_SIZE = 32
    PACKING_BITWIDTH = 32

    def __init__(self, W_nbits=4, group_size=64, in_features=None, out_features=None, input_dtype=DType.FP16,
                 output_dtype=DType.FP16, acc_dtype=None, scaled_activations=False):
        super().__init__()
        if W_nbits not in self.SUPPORTED_BITS_TRITON:
            raise NotImplementedError("Only " + str(self.SUPPORTED_BITS_TRITON) + " W_nbits are supported.")
        if in_features is not None and out_features is not None:
            bad_k = in_features % self.MIN_SIZE != 0
            bad_g = (in_features % group_size != 0) if group_size is not None else False
            if bad_k or bad_g:
                raise NotImplementedError(
                    f"Invalid input shapes: {in_features} , {out_features}. "
                    "in_features should be divisible by 32 or the group_size")
        if input_dtype not in self.SUPPORTED_DTYPES:
            raise NotImplementedError("Unsupport input dtype: " + str(input_dtype))
        if group_size is not None and group_size < 16:
            raise NotImplementedError("Only group_size >= 16 is supported.")

        self.in_features, self.out_features = in_features, out_features
        self.orig_shape = (out_features, in_features)
        self.W_nbits = W_nbits
        self.group_size = 1 if group_size is None else group_size
        self.unpack_mask = 2 ** W_nbits - 1
        self.elements_per_sample = None
        self.signature = (in_features, out_features, W_nbits, self.group_size)
        self.input_dtype, self.output_dtype = input_dtype, output_dtype
        self.compute_dtype = DTYPE_TO_TORCH[input_dtype.value]
        self.meta_dtype = input_dtype
        self.acc_dtype = GEMLITE_ACC_DTYPE[input_dtype] if acc_dtype is None else acc_dtype
        floating = self.compute_dtype in (torch.float16, torch.bfloat16, torch.float32)
        self.scaled_activations = False if floating else scaled_activations
        self.forward = self.forward_auto_no_warmup

    # ---- packing ---------------------------------------------------------------------------------
    def pack(self, W_q: Tensor, scales: Optional[Tensor], zeros: Union[Tensor, int, None], bias: Optional[Tensor] = None,
             fma_mode: bool = True, contiguous: Optional[bool] = None, packing_bitwidth: Optional[int] = None):
        """W_q: (N, K) uint8 codes (bit-packed on the GPU) or int8 / fp8 / fp16 / bf16 weights (kept as a
        (K, N) view).  Resolves W_group_mode / channel_scale_mode exactly like the reference."""
        mx = is_mx_dtype(self.input_dtype)
        if zeros is not None and self.input_dtype == DType.INT8:
            fractional = (isinstance(zeros, Tensor) and zeros.mean() != zeros.int().float().mean()) or \
                isinstance(zeros, float)
            if fractional:
                raise Exception("INT8 inputs is not compatible with floating-point zeros.")
        if packing_bitwidth is None:
            packing_bitwidth = GemLiteLinear.PACKING_BITWIDTH
        if mx:
            packing_bitwidth = 8

        N, K = self.out_features, self.in_features
        self.W_q = None
        if W_q.dtype == torch.int8 or W_q.is_floating_point():
            want = {torch.float32: 32, torch.float16: 16, torch.bfloat16: 16}.get(W_q.dtype, 8)
            assert self.W_nbits == want, f"Invalid {W_q.dtype} weights for W_nbits={self.W_nbits}."
            self.W_q = W_q.t()
            self.elements_per_sample = 1
            if contiguous is None:
                contiguous = False
        elif W_q.dtype == torch.uint8:
            if contiguous is None:
                contiguous = not mx
            codes = W_q.view(N, K)
            if contiguous:   # the kernel writes the (K/eps, N) matrix directly
                self.W_q, self.elements_per_sample = pack_weights_over_cols(codes, self.W_nbits, packing_bitwidth, True)
            else:            # keep the (N, K/eps) storage and expose its transposed view
                packed, self.elements_per_sample = pack_weights_over_cols(codes, self.W_nbits, packing_bitwidth, False)
                self.W_q = packed.t()
        if self.W_q is None:
            raise Exception("Weights were not packed, please check your W_q.dtype")

        self.device = self.W_q.device
        self.bias = None if bias is None else bias.to(device=self.device)

        # ---- mode resolution (core.py:408-464) ----
        group_mode, channel_mode = -1, 0
        if scales is None and zeros is None:
            group_mode, channel_mode = 0, (2 if self.scaled_activations else 0)
        s = scales.view(N, -1).t() if scales is not None else None
        channelwise = (s is not None) and (s.numel() == N)
        self.meta_is_channelwise = channelwise
        z = None
        if zeros is None:
            group_mode = 2 if s is not None else 0
        elif isinstance(zeros, Tensor):
            if fma_mode and not channelwise:      # W ~ fma(Wq, s, -z*s)
                z = (-zeros.float() * scales.float()).to(zeros.dtype).view(N, -1).t()
                group_mode = 4
            else:                                 # W ~ (Wq - z) * s
                z = zeros.view(N, -1).t()
                group_mode = 3
        else:
            z = int(zeros)
            group_mode = 3 if s is not None else 1
        assert group_mode > -1, "Invalid scales/zeros settings."
        if not self.scaled_activations and channelwise:
            channel_mode, group_mode = 1, (1 if z is not None else 0)
        if self.scaled_activations and not channelwise:
            channel_mode = 2
        if self.scaled_activations and channelwise:
            channel_mode, group_mode = 3, (1 if z is not None else 0)
        if channel_mode in (1, 3):
            assert group_mode not in (3, 4), "Can't use channel_scale_mode with W_group_mode == 3 or 4."
        self.W_group_mode, self.channel_scale_mode = group_mode, channel_mode

        # ---- placeholders the functional op can carry (core.py:470-475) ----
        if isinstance(z, int):
            z = torch.tensor(z, dtype=torch.int32, device=self.device)
        if z is None:
            z = torch.tensor([[]], dtype=torch.int32, device=self.device)
        if s is None:
            s = torch.tensor([[]], dtype=torch.int32, device=self.device)

        self.data_contiguous = bool(contiguous)
        if contiguous:
            self.W_q = self.W_q.contiguous()
        s, z = s.contiguous(), z.contiguous()
        if self.input_dtype in (DType.MXFP16, DType.MXBF16, DType.MXFP8, DType.MXFP4):
            s = s.to(torch.float8_e8m0fnu).view(torch.uint8)
        if self.input_dtype == DType.NVFP4:
            s = s.to(torch.float8_e4m3fn)
        if mx:
            s = s.T
            self.W_group_mode, self.channel_scale_mode = 2, 0
        self.meta_dtype = TORCH_TO_DTYPE[s.dtype]

        self.W_q = torch.nn.Parameter(self.W_q, requires_grad=False)
        self.bias = torch.nn.Parameter(self.bias, requires_grad=False) if self.bias is not None else None
        self.scales = torch.nn.Parameter(s, requires_grad=False)
        self.zeros = torch.nn.Parameter(z, requires_grad=False)
        self.metadata = torch.nn.Parameter(
            torch.tensor(self.get_meta_args(), device=self.device, dtype=torch.int32), requires_grad=False)
        self.orig_shape = torch.nn.Parameter(
            torch.tensor([N, K], device=self.device, dtype=torch.int32), requires_grad=False)
        return self

    # ---- (de)serialisation -----------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Accepts the reference's checkpoint layout: W_q, bias, scales, zeros, metadata[12], orig_shape[2]."""
        self.W_q = state_dict.pop("W_q", None)
        self.bias = state_dict.pop("bias", None)
        self.scales = state_dict.pop("scales", None)
        self.zeros = state_dict.pop("zeros", None)
        metadata = state_dict.pop("metadata", None)
        orig_shape = state_dict.pop("orig_shape", None)
        self.metadata = [int(v) for v in metadata.tolist()]
        (self.scaled_activations, self.W_nbits, self.group_size, self.unpack_mask, self.elements_per_sample,
         input_dtype, output_dtype, acc_dtype, meta_dtype, self.channel_scale_mode, self.W_group_mode,
         data_contiguous) = self.metadata
        self.input_dtype, self.output_dtype = DType(input_dtype), DType(output_dtype)
        self.acc_dtype, self.meta_dtype = DType(acc_dtype), DType(meta_dtype)
        self.out_features, self.in_features = [int(v) for v in orig_shape.tolist()]
        self.orig_shape = (self.out_features, self.in_features)
        self.compute_dtype = DTYPE_TO_TORCH[self.input_dtype.value]
        self.scaled_activations = bool(self.scaled_activations)
        self.data_contiguous = bool(data_contiguous)
        self.device = self.W_q.device

    # ---- functional arguments --------------------------------------------------------------------
    def get_tensor_args(self):
        return [self.W_q, self.scales, self.zeros]

    def get_meta_args(self):
        return [int(self.scaled_activations), self.W_nbits, self.group_size, self.unpack_mask,
                self.elements_per_sample, self.input_dtype.value, self.output_dtype.value, self.acc_dtype.value,
                self.meta_dtype.value, self.channel_scale_mode, self.W_group_mode, int(self.data_contiguous)]

    def forward_manual(self, x: Tensor, matmul_type: str = "GEMM") -> Tensor:
        return forward_functional(x, self.bias, self.get_tensor_args(), self.get_meta_args(),
                                  GEMLITE_MATMUL_TYPES_MAPPING[matmul_type])

    def forward_auto_no_warmup(self, x: Tensor) -> Tensor:
        if torch.compiler.is_compiling():      # traced: go through the registered op (fake impl, no graph break)
            return forward_functional(x, self.bias, self.get_tensor_args(), self.get_meta_args())
        return _forward_eager(x, self.bias, (self.W_q, self.scales, self.zeros), self.get_meta_args(), -1)

    # ---- config cache (API kept; there is nothing to autotune on this backend) ---------------------
    @staticmethod
    def cache_config(filename: str):
        try:
            with _FILE_LOCK, open(filename, "r") as f:
                config = json.load(f)
        except Exception:
            config = {}
        for name in GEMLITE_MATMUL_TYPES:
            config.setdefault(name, {})
            config[name].update(GEMLITE_TRITON_CONFIG_CACHE.get(name, {}))
        with _FILE_LOCK, open(filename, "w") as f:
            json.dump(config, f)

    @staticmethod
    def load_config(filename: str, print_error: bool = True, overwrite: bool = False):
        global GEMLITE_TRITON_CONFIG_CACHE
        if filename is None:
            return False
        try:
            with _FILE_LOCK, open(filename, "r") as f:
                config = json.load(f)
            if overwrite:
                GEMLITE_TRITON_CONFIG_CACHE = config
            else:
                for name, entries in config.items():
                    GEMLITE_TRITON_CONFIG_CACHE.setdefault(name, {}).update(entries)
        except Exception as e:
            if print_error:
                logger.error(f"Failed to load the cache file '{filename}': {e}")
            return False
        return True

    @staticmethod
    def reset_config():
        global GEMLITE_TRITON_CONFIG_CACHE
        GEMLITE_TRITON_CONFIG_CACHE = {}


GemLiteLinearTriton = GemLiteLinear   # the reference's class name (callers import either)
GemLiteLinearCUDA = GemLiteLinear
