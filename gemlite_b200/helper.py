"""Quantise-and-pack "processor" classes and model patching, API-compatible with gemlite/helper.py.

Each processor turns weights (or a torch.nn.Linear / HQQLinear / BitLinear) into a packed
`GemLiteLinear`.  What they configure -- dtypes, group size, scaled activations and the
(W_group_mode, channel_scale_mode) override applied after pack() -- follows the reference recipe
cited on each class; the packing and all compute run in the sm_100a CUDA backend.
"""
import gc

import torch

from .core import GemLiteLinear, GemLiteLinearTriton  # noqa: F401
from .dtypes import DType, TORCH_TO_DTYPE
from .quant_utils import WeightQuantizerMXFP

default_fp8 = torch.float8_e4m3fn
default_post_scale = False
_FLOATS = (torch.float16, torch.bfloat16, torch.float32)


# --------------------------------------------------------------------------------------------------
# small shared pieces
# --------------------------------------------------------------------------------------------------
def cleanup_linear(linear_layer, del_orig=True):
    if del_orig:
        for attr in ("weight", "bias", "weight_scale", "W_q", "meta"):
            val = getattr(linear_layer, attr, None)
            if val is not None and hasattr(val, "__len__") and len(val) > 0:
                setattr(linear_layer, attr, None)
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def _data(t):
    return t.data if isinstance(t, torch.nn.Parameter) else t


def _value_range(dtype):
    info = torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)
    return info.min, info.max


def _quantize_per_channel(weight, w_dtype, device):
    """fp32 absmax per output row / qmax (clamped at 1e-6); int8 rounds half-to-even, fp8 casts.
    (helper.py:119-128, 436-444)"""
    lo, hi = _value_range(w_dtype)
    w32 = weight.to(dtype=torch.float32, device=device)
    scales = w32.abs().amax(dim=1, keepdim=True) / hi
    scales.clamp_(min=1e-6)
    q = (w32 / scales).clamp_(min=lo, max=hi)
    q = q.to(w_dtype) if w_dtype.is_floating_point else q.round_().to(w_dtype)
    return q, scales


def _set_modes(layer, group_mode, channel_mode):
    layer.W_group_mode, layer.channel_scale_mode = group_mode, channel_mode
    return layer


def _hqq_parts(hqq_layer):
    """(W_q uint8 (N,K), scales, zeros, bias, nbits, group_size) of an hqq.core.quantize.HQQLinear (axis 1)."""
    assert hqq_layer.meta["axis"] == 1, "Only axis==1 is supported."
    nbits = hqq_layer.meta["nbits"]
    group_size = hqq_layer.meta["group_size"] or hqq_layer.in_features
    W_q = hqq_layer.unpack(dtype=torch.uint8).view(hqq_layer.meta["shape"])
    scales, zeros = hqq_layer.meta["scale"].clone(), hqq_layer.meta["zero"].clone()
    bias = hqq_layer.bias.clone() if hqq_layer.bias is not None else None
    return W_q, scales, zeros, bias, nbits, group_size


# --------------------------------------------------------------------------------------------------
# model patching / warm-up
# --------------------------------------------------------------------------------------------------
def patch_model(model, device, processor, skip_modules=["lm_head", "vision", "visual"], group_size=64):
    """Replace every torch.nn.Linear (or HQQLinear) whose qualified name does not contain one of
    `skip_modules` by the processor's packed layer. `group_size` only matters for HQQ (W_nbits <= 4)."""
    hqq_cls, make_cfg = (), None
    if hasattr(processor, "from_hqqlinear"):
        try:
            from hqq.core.quantize import HQQLinear, BaseQuantizeConfig
            hqq_cls, make_cfg = (HQQLinear,), BaseQuantizeConfig
        except ImportError:
            print("This processor requires the `hqq` package. Install it via `pip install hqq`.")

    for name, module in model.named_modules():
        module.name = name

    def convert(layer):
        layer = layer.to(device=device, non_blocking=True)
        if any(s in layer.name for s in skip_modules):
            return layer
        if hasattr(processor, "from_hqqlinear") and make_cfg is not None:
            if isinstance(layer, torch.nn.Linear):
                nbits = processor.W_nbits
                cfg = make_cfg(nbits=nbits, group_size=group_size if nbits <= 4 else None)
                layer = hqq_cls[0](layer, quant_config=cfg, compute_dtype=layer.weight.dtype, device=device)
            return processor.from_hqqlinear(layer)
        return processor.from_linear(layer)

    def walk(parent):
        for name, child in parent.named_children():
            if isinstance(child, (torch.nn.Linear,) + hqq_cls):
                setattr(parent, name, convert(child))
            else:
                walk(child)

    walk(model)
    model = model.to(device=device)
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    gc.collect()
    return model


default_batch_sizes = [4096, 2048, 1024, 512, 256, 128, 64, 32, 16, 8, 4, 2, 1]


def warmup(processor, shapes, batch_sizes=default_batch_sizes, group_size=64, dtype=torch.float16):
    """Run each (out_features, in_features) shape once per batch size. There is nothing to autotune on
    this backend; this only pays one-time costs (module load, shared-memory opt-in, workspaces)."""
    device = torch.device(torch.cuda.current_device())
    for out_features, in_features in shapes:
        linear = torch.nn.Linear(in_features, out_features, bias=False, dtype=dtype, device=device)
        if hasattr(processor, "from_hqqlinear") and not hasattr(processor, "from_linear"):
            from hqq.core.quantize import HQQLinear, BaseQuantizeConfig
            cfg = BaseQuantizeConfig(nbits=processor.W_nbits, group_size=group_size, axis=1)
            layer = processor.from_hqqlinear(HQQLinear(linear, quant_config=cfg, compute_dtype=dtype, device=device))
        else:
            layer = processor.from_linear(linear)
        for bs in batch_sizes:
            layer(torch.randn((bs, in_features), dtype=dtype, device=device) / 100.0)
        torch.cuda.synchronize()
        del layer
        gc.collect()


# --------------------------------------------------------------------------------------------------
# A16W8: 16-bit activations, 8-bit (int8 / fp8) channel-wise weights            (helper.py:88-170)
# --------------------------------------------------------------------------------------------------
class A16W8:
    def __init__(self, device="cuda:0", dtype=None, fp8=None, fp32_scale=True, post_scale=False):
        self.device, self.dtype, self.fp8 = device, dtype, fp8
        self.fp32_scale, self.post_scale = fp32_scale, post_scale

    def from_weights(self, weight, bias=None, scales=None):
        weight, bias = _data(weight), _data(bias)
        out_features, in_features = weight.shape
        if scales is None:
            dtype = weight.dtype if self.dtype is None else self.dtype
            assert dtype in _FLOATS, f"Invalid weight dtype, should be floating point, got {dtype}"
            W_q, scales = _quantize_per_channel(weight, self.fp8 if self.fp8 else torch.int8, self.device)
        else:
            assert weight.itemsize == 1, f"Invalid weight.dtype, should be 8-bit (INT8 or FP8), got {weight.dtype}"
            dtype = self.dtype if self.dtype is not None else \
                (scales.dtype if scales.dtype in (torch.float16, torch.bfloat16) else torch.float16)
            W_q, scales = weight.to(device=self.device), scales.to(device=self.device)
        gd = TORCH_TO_DTYPE[dtype]
        bias = bias.to(device=self.device, dtype=dtype) if bias is not None else None
        layer = GemLiteLinear(8, group_size=in_features, in_features=in_features, out_features=out_features,
                              input_dtype=gd, output_dtype=gd)
        layer.pack(W_q, scales, zeros=None, bias=bias)
        # post_scale: scale the accumulator per output channel; otherwise dequantise before the product
        return _set_modes(layer, 0, 1) if self.post_scale else _set_modes(layer, 2, 0)

    def from_linear(self, linear_layer, del_orig=True):
        out = self.from_weights(weight=linear_layer.weight, bias=linear_layer.bias)
        cleanup_linear(linear_layer, del_orig)
        return out


class A16W8_INT8(A16W8):
    def __init__(self, device="cuda:0", dtype=None, fp32_scale=True, post_scale=False):
        super().__init__(device=device, dtype=dtype, fp8=None, fp32_scale=fp32_scale, post_scale=post_scale)


class A16W8_FP8(A16W8):
    def __init__(self, device="cuda:0", dtype=None, fp8=default_fp8, fp32_scale=True, post_scale=False):
        super().__init__(device=device, dtype=dtype, fp8=fp8, fp32_scale=fp32_scale, post_scale=post_scale)


# --------------------------------------------------------------------------------------------------
# A16Wn: 16-bit activations, packed 8/4/2/1-bit grouped INT or 8/4-bit MXFP weights (helper.py:172-400)
# --------------------------------------------------------------------------------------------------
class A16Wn:
    def __init__(self, device="cuda:0", dtype=None, packing_bitwidth=None, post_scale=default_post_scale):
        self.device, self.dtype = device, dtype
        self.packing_bitwidth, self.post_scale = packing_bitwidth, post_scale
        self.quantizer_mx = None
        self.mx_fp8_dtype = default_fp8

    def from_weights(self, W_q, scales, zeros, W_nbits, group_size, bias=None, quant_type="INT"):
        return self.from_weights_(W_q, scales, zeros, W_nbits, group_size, bias, quant_type)

    def from_weights_(self, W_q, scales, zeros, W_nbits, group_size, bias=None, quant_type="INT"):
        assert quant_type in ["INT", "MXFP"], f"Invalid quant_type. Got {quant_type}, valid values are INT, MXFP."
        W_q, bias = _data(W_q), _data(bias)
        if quant_type == "INT":
            dtype = scales.dtype if self.dtype is None else self.dtype
            assert dtype in _FLOATS, "Invalid scales.dtype, should floating point."
            gd = TORCH_TO_DTYPE[dtype]
            scales, zeros = scales.to(dtype), zeros.to(dtype)
        else:
            assert W_nbits in [8, 4], "Unsupported W_nbit for MXFP quant_dtype."
            assert group_size == 32, "group_size should 32 for MXFP."
            want = self.mx_fp8_dtype if W_nbits == 8 else torch.uint8
            assert W_q.dtype == want, f"Unsupported dtype of W_q. got {W_q.dtype}"
            dtype = torch.float16 if self.dtype is None else self.dtype
            if dtype not in (torch.float16, torch.bfloat16):
                raise Exception(f"Unsupported dtype for MXFP. Got {dtype}, supported [torch.float16, torch.bfloat16]")
            gd = DType.MXFP16 if dtype == torch.float16 else DType.MXBF16
            self.post_scale = False
            N, K = W_q.shape
            scales = scales.view(N, K // group_size)
        out_features, in_features = W_q.shape
        W_q = W_q.to(self.device)
        scales = scales.to(device=self.device) if scales is not None else None
        zeros = zeros.to(device=self.device) if zeros is not None else None
        bias = bias.to(device=self.device, dtype=dtype) if bias is not None else None

        layer = GemLiteLinear(W_nbits, group_size=group_size, in_features=in_features, out_features=out_features,
                              input_dtype=gd, output_dtype=gd, scaled_activations=False)
        layer.pack(W_q, scales, zeros, bias=bias, packing_bitwidth=self.packing_bitwidth)
        if group_size == in_features and quant_type == "INT":     # one group per row
            _set_modes(layer, 1, 1) if self.post_scale else _set_modes(layer, 3, 0)
        return layer

    def from_hqqlinear(self, hqq_layer, del_orig=True):
        self.device = hqq_layer.W_q.device
        W_q, scales, zeros, bias, nbits, group_size = _hqq_parts(hqq_layer)
        cleanup_linear(hqq_layer, del_orig)
        return self.from_weights_(W_q=W_q, scales=scales, zeros=zeros, W_nbits=nbits, group_size=group_size,
                                  bias=bias, quant_type="INT")

    def mxfp_from_linear(self, linear_layer, W_nbits, del_orig=True):
        if self.quantizer_mx is None:
            self.quantizer_mx = WeightQuantizerMXFP(device=self.device, compute_dtype=linear_layer.weight.dtype)
        W = linear_layer.weight.data.to(self.device)
        bias = linear_layer.bias.clone() if linear_layer.bias is not None else None
        N, K = W.shape
        if W_nbits == 8:
            W_q, scales = self.quantizer_mx.quantize_mxfp8(W, index=True, mx_fp8_dtype=self.mx_fp8_dtype)
        else:
            W_q, scales = self.quantizer_mx.quantize_mxfp4(W, index=True)
        cleanup_linear(linear_layer, del_orig)
        return self.from_weights_(W_q=W_q.view(N, K), scales=scales.view(N, K // 32), zeros=None, W_nbits=W_nbits,
                                  group_size=32, bias=bias, quant_type="MXFP")


class A16Wn_HQQ_INT(A16Wn):
    def __init__(self, device="cuda:0", dtype=None, W_nbits=None):
        super().__init__(device=device, dtype=dtype)
        self.W_nbits = W_nbits

    def from_weights(self, W_q, scales, zeros, bias=None):
        group_size = W_q.numel() // scales.numel()
        return super().from_weights(W_q=W_q, scales=scales, zeros=zeros, W_nbits=self.W_nbits,
                                    group_size=group_size, bias=bias, quant_type="INT")


class A16W8_HQQ_INT(A16Wn_HQQ_INT):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, W_nbits=8)


class A16W4_HQQ_INT(A16Wn_HQQ_INT):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, W_nbits=4)


class A16W2_HQQ_INT(A16Wn_HQQ_INT):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, W_nbits=2)


class A16W1_HQQ_INT(A16Wn_HQQ_INT):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, W_nbits=1)


class A16Wn_MXFP(A16Wn):
    def __init__(self, device="cuda:0", dtype=None, W_nbits=None):
        super().__init__(device=device, dtype=dtype)
        self.W_nbits = W_nbits

    def from_weights(self, W_q, scales, bias=None):
        group_size = W_q.numel() // scales.numel()
        return super().from_weights(W_q=W_q, scales=scales, zeros=None, W_nbits=self.W_nbits,
                                    group_size=group_size, bias=bias, quant_type="MXFP")

    def from_linear(self, linear_layer, del_orig=True):
        return super().mxfp_from_linear(linear_layer=linear_layer, W_nbits=self.W_nbits, del_orig=del_orig)


class A16W8_MXFP(A16Wn_MXFP):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, W_nbits=8)


class A16W4_MXFP(A16Wn_MXFP):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, W_nbits=4)


# --------------------------------------------------------------------------------------------------
# A8W8 dynamic: per-token int8 / fp8 activations x channel-wise 8-bit weights     (helper.py:405-497)
# --------------------------------------------------------------------------------------------------
class A8W8_dynamic:
    def __init__(self, device="cuda:0", dtype=None, fp8=False, fp32_scale=True):
        self.device, self.dtype, self.fp8, self.fp32_scale = device, dtype, fp8, fp32_scale

    def from_weights(self, weight, bias=None, scales=None):
        weight, bias = _data(weight), _data(bias)
        w_dtype = self.fp8 if self.fp8 else torch.int8
        input_dtype = TORCH_TO_DTYPE[w_dtype]
        out_features, in_features = weight.shape
        if scales is None:
            dtype = weight.dtype if self.dtype is None else self.dtype
            assert dtype in _FLOATS, "Invalid weight dtype, should be floating point."
            W_q, scales = _quantize_per_channel(weight, w_dtype, self.device)
        else:
            assert weight.dtype.itemsize == 1, "Invalid weight.dtype, should be 8-bit."
            dtype = self.dtype if self.dtype is not None else \
                (scales.dtype if scales.dtype in (torch.float16, torch.bfloat16) else torch.float16)
            W_q, scales = weight.to(device=self.device), scales.to(device=self.device)
        scales = scales.to(torch.float32 if self.fp32_scale else dtype)
        bias = bias.to(device=self.device, dtype=dtype) if bias is not None else None
        layer = GemLiteLinear(8, group_size=in_features, in_features=in_features, out_features=out_features,
                              input_dtype=input_dtype, output_dtype=TORCH_TO_DTYPE[dtype], scaled_activations=True)
        layer.pack(W_q, scales, zeros=None, bias=bias)
        return _set_modes(layer, 0, 3)       # both scales applied to the accumulator

    def from_linear(self, linear_layer, del_orig=True):
        out = self.from_weights(weight=linear_layer.weight, bias=linear_layer.bias)
        cleanup_linear(linear_layer, del_orig)
        return out


class A8W8_int8_dynamic(A8W8_dynamic):
    def __init__(self, device="cuda:0", dtype=None):
        super().__init__(device=device, dtype=dtype, fp8=False)


class A8W8_fp8_dynamic(A8W8_dynamic):
    def __init__(self, device="cuda:0", dtype=None, fp8=default_fp8):
        super().__init__(device=device, dtype=dtype, fp8=fp8)


A8W8_INT8_dynamic = A8W8_int8_dynamic
A8W8_FP8_dynamic = A8W8_fp8_dynamic


# --------------------------------------------------------------------------------------------------
# A8Wn dynamic: per-token fp8 activations x packed grouped INT weights            (helper.py:502-655)
# --------------------------------------------------------------------------------------------------
class A8Wn_HQQ_INT_dynamic(A16Wn):
    def __init__(self, device="cuda:0", packing_bitwidth=None, dtype=None, post_scale=default_post_scale,
                 fp8=default_fp8, fp32_scale=False, W_nbits=None):
        assert W_nbits is not None, "W_nbits argument should be eitehr 8 or 4, not None)."
        super().__init__(device=device, dtype=dtype, packing_bitwidth=packing_bitwidth, post_scale=post_scale)
        self.fp8, self.fp32_scale, self.W_nbits = fp8, fp32_scale, W_nbits

    def from_weights(self, W_q, scales, zeros, bias=None):
        group_size = W_q.numel() // scales.numel()
        return self.from_weights_(W_q=W_q, scales=scales, zeros=zeros, W_nbits=self.W_nbits,
                                  group_size=group_size, bias=bias)

    def from_weights_(self, W_q, scales, zeros, W_nbits, group_size, bias=None):
        W_q, bias = _data(W_q), _data(bias)
        dtype = self.dtype if self.dtype is not None else \
            (scales.dtype if scales.dtype in (torch.float16, torch.bfloat16) else torch.float16)
        assert dtype in _FLOATS, "Invalid scales.dtype, should be floating point."
        W_q = W_q.to(self.device)
        scales = scales.to(device=self.device, dtype=torch.float32 if self.fp32_scale else dtype)
        zeros = zeros.to(device=self.device, dtype=dtype) if zeros is not None else None
        bias = bias.to(device=self.device, dtype=dtype) if bias is not None else None
        out_features, in_features = W_q.shape
        layer = GemLiteLinear(W_nbits, group_size=group_size, in_features=in_features, out_features=out_features,
                              input_dtype=TORCH_TO_DTYPE[self.fp8], output_dtype=TORCH_TO_DTYPE[dtype],
                              scaled_activations=True)
        layer.pack(W_q, scales, zeros, bias=bias, packing_bitwidth=self.packing_bitwidth, fma_mode=False)
        if group_size == in_features:
            _set_modes(layer, 1, 3) if self.post_scale else _set_modes(layer, 3, 2)
        return layer

    def from_hqqlinear(self, hqq_layer, del_orig=True):
        self.device = hqq_layer.W_q.device
        W_q, scales, zeros, bias, nbits, group_size = _hqq_parts(hqq_layer)
        cleanup_linear(hqq_layer, del_orig)
        return self.from_weights_(W_q=W_q, scales=scales, zeros=zeros, W_nbits=nbits, group_size=group_size, bias=bias)


class A8W4_HQQ_INT_dynamic(A8Wn_HQQ_INT_dynamic):
    def __init__(self, device="cuda:0", packing_bitwidth=None, dtype=None, post_scale=default_post_scale,
                 fp8=default_fp8, fp32_scale=False):
        super().__init__(device=device, packing_bitwidth=packing_bitwidth, dtype=dtype, post_scale=post_scale,
                         fp8=fp8, fp32_scale=fp32_scale, W_nbits=4)


class A8W2_HQQ_INT_dynamic(A8Wn_HQQ_INT_dynamic):
    def __init__(self, device="cuda:0", packing_bitwidth=None, dtype=None, post_scale=default_post_scale,
                 fp8=default_fp8, fp32_scale=False):
        super().__init__(device=device, packing_bitwidth=packing_bitwidth, dtype=dtype, post_scale=post_scale,
                         fp8=fp8, fp32_scale=fp32_scale, W_nbits=2)


# --------------------------------------------------------------------------------------------------
# MXFP / NVFP weights with dynamically quantised activations                      (helper.py:660-945)
# (packing is supported; the activation-quantisation kernels for these are the next scope item)
# --------------------------------------------------------------------------------------------------
class _MXDynamicBase:
    W_nbits = 8
    group_size = 32
    input_dtype = DType.MXFP8

    def __init__(self, device="cuda:0", dtype=None):
        self.device, self.dtype = device, dtype
        self.quantizer_mx = None
        self.mx_fp8_dtype = default_fp8

    def _modes(self):
        return 0, 4

    def _check_inputs(self, weight, scales):
        if self.W_nbits == 8:
            assert weight.is_floating_point() and weight.itemsize == 1, \
                f"Invalid weight.dtype, should be an MXPF8 valid dtype, got {weight.dtype}."
        else:
            assert weight.dtype == torch.uint8, f"Invalid weight.dtype, should be uint8 fp4 codes, got {weight.dtype}."
        ok = (torch.float8_e4m3fn,) if self.input_dtype == DType.NVFP4 else (torch.float8_e8m0fnu, torch.uint8)
        assert scales.dtype in ok, f"Invalid scales.dtype {scales.dtype}."

    def from_weights(self, weight, bias=None, scales=None):
        weight, bias = _data(weight), _data(bias)
        assert scales is not None, "Scales parameter cannot be None. Use from_linear() call to pre-quantize the weights."
        assert self.dtype is not None, "Input dtype should be either torch.float16 or torch.bfloat16, not None."
        self._check_inputs(weight, scales)
        out_features, in_features = weight.shape
        bias = bias.to(device=self.device, dtype=self.dtype) if bias is not None else None
        layer = GemLiteLinear(self.W_nbits, group_size=self.group_size, in_features=in_features,
                              out_features=out_features, input_dtype=self.input_dtype,
                              output_dtype=TORCH_TO_DTYPE[self.dtype], scaled_activations=True)
        layer.pack(weight.to(device=self.device), scales.to(device=self.device), zeros=None, bias=bias)
        return _set_modes(layer, *self._modes())

    def _quantize(self, W):
        if self.input_dtype == DType.NVFP4:
            return self.quantizer_mx.quantize_nvfp4(W, index=True)
        if self.W_nbits == 8:
            return self.quantizer_mx.quantize_mxfp8(W, index=True, mx_fp8_dtype=self.mx_fp8_dtype)
        return self.quantizer_mx.quantize_mxfp4(W, index=True)

    def from_linear(self, linear_layer, del_orig=True):
        if self.quantizer_mx is None:
            self.quantizer_mx = WeightQuantizerMXFP(device=self.device, compute_dtype=linear_layer.weight.dtype)
        W = linear_layer.weight.data.to(self.device)
        bias = linear_layer.bias.clone() if linear_layer.bias is not None else None
        N, K = W.shape
        W_q, scales = self._quantize(W)
        cleanup_linear(linear_layer, del_orig)
        return self.from_weights(weight=W_q.view(N, K), scales=scales.view(N, K // self.group_size), bias=bias)


class A8Wn_MXFP_dynamic(_MXDynamicBase):
    def __init__(self, device="cuda:0", dtype=None, post_scale=True, fp8=default_fp8, W_nbits=None):
        assert W_nbits is not None, "W_nbits argument should be eitehr 8 or 4, not None)."
        super().__init__(device=device, dtype=dtype)
        self.mx_fp8_dtype, self.post_scale, self.W_nbits = fp8, post_scale, W_nbits

    def _modes(self):   # per-token post-scale (2) or mxfp8 micro-scales for the activations (4)
        return 0, (2 if self.post_scale else 4)


class A8W8_MXFP_dynamic(A8Wn_MXFP_dynamic):
    def __init__(self, device="cuda:0", dtype=None, post_scale=True, fp8=default_fp8):
        super().__init__(device=device, dtype=dtype, post_scale=post_scale, fp8=fp8, W_nbits=8)


class A8W4_MXFP_dynamic(A8Wn_MXFP_dynamic):
    def __init__(self, device="cuda:0", dtype=None, post_scale=True, fp8=default_fp8):
        super().__init__(device=device, dtype=dtype, post_scale=post_scale, fp8=fp8, W_nbits=4)


class A4W4_MXFP_dynamic(_MXDynamicBase):
    W_nbits, group_size, input_dtype = 4, 32, DType.MXFP4


class A4W4_NVFP_dynamic(_MXDynamicBase):
    W_nbits, group_size, input_dtype = 4, 16, DType.NVFP4


# --------------------------------------------------------------------------------------------------
# BitNet (ternary weights stored as 2-bit codes w + 1)                            (helper.py:950-1062)
# --------------------------------------------------------------------------------------------------
class A16W158_INT:
    _input_dtype = None
    _scaled_activations = False
    _modes = (1, 1)      # shift by the integer zero-point, weight scale on the accumulator

    def __init__(self, device="cuda:0", dtype=None, fp32_scale=True):
        self.device, self.dtype, self.fp32_scale = device, dtype, fp32_scale

    def from_weights(self, weight, weight_scale, bias=None):
        weight, bias = _data(weight), _data(bias)
        dtype = weight.dtype if self.dtype is None else self.dtype
        assert dtype in _FLOATS, "Invalid weight.dtype, should be floating point."
        weight = weight.to(device=self.device, dtype=dtype)
        W_q = (weight + 1).to(torch.uint8)
        out_features, in_features = W_q.shape
        gd = TORCH_TO_DTYPE[dtype]
        scales = torch.full((out_features, 1), float(weight_scale.item()), dtype=torch.float32, device=self.device)
        scales = scales.to(dtype=torch.float32 if self.fp32_scale else dtype)
        bias = bias.to(device=self.device, dtype=dtype) if bias is not None else None
        layer = GemLiteLinear(2, group_size=in_features, in_features=in_features, out_features=out_features,
                              input_dtype=self._input_dtype or gd, output_dtype=gd,
                              scaled_activations=self._scaled_activations)
        layer.pack(W_q, scales=scales, zeros=1, bias=bias)
        return _set_modes(layer, *self._modes)

    def from_bitlinear(self, linear_layer, del_orig=True):
        out = self.from_weights(weight=linear_layer.weight, weight_scale=linear_layer.weight_scale,
                                bias=linear_layer.bias)
        cleanup_linear(linear_layer, del_orig)
        return out


class A8W158_INT_dynamic(A16W158_INT):
    _input_dtype = DType.INT8
    _scaled_activations = True
    _modes = (1, 3)      # activation and weight scales on the accumulator
