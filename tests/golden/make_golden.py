"""Generate golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz + MANIFEST.json

The reference (/root/reference, read-only) is a Triton library whose import touches the GPU.
There is no GPU here, so this script
  * answers the import-time device probes with a fake "NVIDIA B200" record,
  * installs a stub Triton driver and sets TRITON_INTERPRET=1, so that every `@triton.jit`
    kernel of the reference is *executed by Triton's CPU interpreter* (numpy), and
  * works around one interpreter bug (python `bool` launch arguments are mangled as int1 but
    materialised as int32) without touching reference code, and
  * supplies numpy stand-ins for the two primitives the interpreter lacks: `libdevice.round` and `tl.dot_scaled`
    (see the comments at their definitions; everything around them is the reference's code), and
  * replaces the interpreter's float down-casts (fp32 -> bf16 truncates, fp32 -> fp8 mis-rounds) with the
    round-to-nearest-even (fp8: saturating) conversions the GPU performs.
Everything that is recorded in the fixtures is produced by reference code: `pack()`,
`forward_manual()` of each kernel that the interpreter can execute, `bitpack`, the
per-token activation quantiser and the MXFP weight quantiser.

The fixtures cannot be regenerated on the GPU box (no /root/reference there); they are small
and committed.  Floating tensors are stored as raw bit patterns (uint16/uint8 views) so that
.npz round-trips are exact.
"""
import json
import os
import sys
import traceback

os.environ["TRITON_INTERPRET"] = "1"
import numpy as np
import torch
import torch._dynamo as _dynamo

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# ---------------------------------------------------------------------------------------------
# environment shims (no reference code is modified)
# ---------------------------------------------------------------------------------------------
class _Props:
    name = "NVIDIA B200"
    multi_processor_count = 148
    major, minor = 10, 0
    total_memory = 180 << 30


torch.cuda.get_device_properties = lambda *a, **k: _Props()
torch.cuda.device_count = lambda: 0
torch.cuda.empty_cache = lambda: None

import triton  # noqa: E402
import triton.language as tl  # noqa: E402
import triton.runtime.interpreter as _interp  # noqa: E402
from triton.backends.compiler import GPUTarget  # noqa: E402
from triton.backends.driver import DriverBase  # noqa: E402


class _StubDriver(DriverBase):
    def __init__(self):
        self.get_current_stream = lambda *a: 0
        self.get_current_device = lambda *a: 0
        self.set_current_device = lambda *a: None
        self.get_device_capability = lambda *a: (10, 0)

    @classmethod
    def is_active(cls):
        return True

    def get_current_target(self):
        return GPUTarget("cuda", 100, 32)

    def get_active_torch_device(self):
        return torch.device("cpu")

    def get_device_interface(self):
        return torch.cuda

    def get_benchmarker(self):
        return lambda *a, **k: [0.0, 0.0, 0.0]

    def get_empty_cache_for_benchmark(self):
        return torch.empty(16)

    def clear_cache(self, c):
        pass

    def map_python_to_cpp_type(self, ty):
        return ty


triton.runtime.driver.set_active(_StubDriver())

# The interpreter's float DOWN-casts are wrong for the types stored as raw bits: fp32 -> bf16 truncates, and
# fp32 -> fp8 "rounds" by adding the cut-off bit without carrying into the exponent (0.485 -> 0.25, 31 -> 16, ties go
# up).  The GPU converts with cvt.rn(.satfinite): round-to-nearest-even, fp8 saturating.  Replace the down-casts with
# exactly that (torch's conversions are RNE); up-casts, which are exact, keep the interpreter's code.
_TL2TORCH = {tl.float32: torch.float32, tl.float16: torch.float16, tl.bfloat16: torch.bfloat16,
             tl.float8e4nv: torch.float8_e4m3fn, tl.float8e5: torch.float8_e5m2}
_BITS_VIEW = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
_orig_convert_float = _interp._convert_float


def _convert_float_rne(input, input_dtype, output_dtype, rounding_mode):
    if (input_dtype in _TL2TORCH and output_dtype in _TL2TORCH
            and input_dtype.primitive_bitwidth > output_dtype.primitive_bitwidth):
        raw = np.ascontiguousarray(input)
        ti, to = _TL2TORCH[input_dtype], _TL2TORCH[output_dtype]
        src = torch.from_numpy(raw.view(getattr(np, f"uint{raw.dtype.itemsize * 8}")).astype(
            {1: np.uint8, 2: np.int16, 4: np.int32}[raw.dtype.itemsize], copy=False)).view(ti).to(torch.float32)
        if to in (torch.float8_e4m3fn, torch.float8_e5m2):
            hi = torch.finfo(to).max
            src = torch.where(torch.isfinite(src), src.clamp(-hi, hi), src)          # .satfinite
        out = src.to(to)
        out_bits = out.view(_BITS_VIEW[out.element_size()]).numpy()
        return out_bits.view(getattr(np, f"uint{out.element_size() * 8}")).reshape(raw.shape)
    return _orig_convert_float(input, input_dtype, output_dtype, rounding_mode)


_interp._convert_float = _convert_float_rne

_orig_cvt = _interp._implicit_cvt


def _cvt(arg):
    if isinstance(arg, bool):
        return tl.tensor(_interp.TensorHandle(np.array([arg], dtype=np.bool_), tl.int1), tl.int1)
    return _orig_cvt(arg)


_interp._implicit_cvt = _cvt

sys.path.insert(0, REF)
import gemlite  # noqa: E402
from gemlite import bitpack as ref_bitpack  # noqa: E402
from gemlite import quant_utils as ref_qu  # noqa: E402
from gemlite.core import GemLiteLinearTriton, DType, GEMLITE_MATMUL_TYPES_MAPPING, forward_functional  # noqa: E402
from gemlite.triton_kernels import gemv_kernels as _gk  # noqa: E402

gemlite.reset_config()
gemlite.set_autotune(False)


# libdevice (a CUDA bitcode library) does not exist for the CPU interpreter: give the activation
# quantiser an exact restatement of libdevice.round (round half away from zero) to call instead.
@triton.jit
def _round_half_away(x):
    r = tl.where(x >= 0, tl.floor(x), tl.ceil(x))
    d = tl.abs(x - r)
    return tl.where(d >= 0.5, r + tl.where(x >= 0, 1.0, -1.0), r)


class _LibdeviceShim:
    round = _round_half_away


ref_qu.libdevice = _LibdeviceShim


# tl.dot_scaled is not implemented by the CPU interpreter (its builder has no create_dot_scaled).  The MX GEMM kernels
# of the reference are therefore run with a numpy stand-in for that ONE compiler primitive, following its documented
# semantics: operands decoded exactly (e2m1 low nibble = even k, e4m3 / e5m2 / fp16 / bf16 bit patterns), each
# 32- (or 16-) element group of K multiplied by its scale (uint8 = e8m0 -> 2^(e-127), float8 = e4m3 value), products
# summed exactly (float64) and rounded once into the fp32 accumulator.  Everything around the primitive -- pointer
# arithmetic, scale indexing, masks, the NVFP4 norm, the post-scale epilogue -- is the reference's own code.
_FP4_TABLE = np.array([0, 0.5, 1, 1.5, 2, 3, 4, 6, -0.0, -0.5, -1, -1.5, -2, -3, -4, -6], dtype=np.float64)


def _ds_decode(t, fmt, k_axis):
    d = t.handle.data
    if fmt == "e2m1":
        d = d.astype(np.uint8)
        st = np.stack([d & 15, d >> 4], axis=k_axis + 1)
        shape = list(d.shape)
        shape[k_axis] *= 2
        return _FP4_TABLE[st.reshape(shape)]
    if fmt == "e4m3":
        return torch.from_numpy(d.view(np.uint8).copy()).view(torch.float8_e4m3fn).to(torch.float64).numpy()
    if fmt == "e5m2":
        return torch.from_numpy(d.view(np.uint8).copy()).view(torch.float8_e5m2).to(torch.float64).numpy()
    if fmt == "bf16":
        return torch.from_numpy(d.view(np.int16).copy()).view(torch.bfloat16).to(torch.float64).numpy()
    if fmt == "fp16":
        return d.astype(np.float64)
    raise NotImplementedError(fmt)


def _ds_scale(t):
    d = t.handle.data
    if t.dtype.scalar == tl.uint8:
        return np.exp2(d.astype(np.float64) - 127.0)
    return torch.from_numpy(d.view(np.uint8).copy()).view(torch.float8_e4m3fn).to(torch.float64).numpy()


def _dot_scaled_standin(lhs, lhs_scale, lhs_format, rhs, rhs_scale, rhs_format, acc=None, **kw):
    lf, rf = getattr(lhs_format, "value", lhs_format), getattr(rhs_format, "value", rhs_format)
    A = _ds_decode(lhs, lf, 1)                      # (M, K)
    B = _ds_decode(rhs, rf, 0)                      # (K, N)
    K = A.shape[1]
    assert B.shape[0] == K
    if lhs_scale is not None and getattr(lhs_scale, "value", 0) is not None:
        sa = _ds_scale(lhs_scale)                   # (M, K / gs)
        A = A * np.repeat(sa, K // sa.shape[1], axis=1)
    if rhs_scale is not None and getattr(rhs_scale, "value", 0) is not None:
        sb = _ds_scale(rhs_scale)                   # (N, K / gs)
        B = B * np.repeat(sb, K // sb.shape[1], axis=1).T
    out = (A @ B).astype(np.float32)
    if acc is not None:
        out = out + acc.handle.data.astype(np.float32)
    return tl.tensor(_interp.TensorHandle(out.astype(np.float32), tl.float32), tl.block_type(tl.float32, list(out.shape)))


tl.dot_scaled = _dot_scaled_standin


class _L(list):  # per-device LUT list indexed with device.index (None on CPU)
    def __getitem__(self, i):
        return list.__getitem__(self, 0 if i is None else i)


_gk.fp4_mapping = _L([torch.tensor([0, 1, 2, 3, 4, 6, 8, 12, 0, -1, -2, -3, -4, -6, -8, -12], dtype=torch.int8)])
for _name in ("fp4_values", "fp4_p_vals", "fp4_thresholds", "thr_pos"):
    _lst = getattr(ref_qu, _name)
    _vals = {
        "fp4_values": [0, 0.5, 1, 1.5, 2, 3, 4, 6, -0.0, -0.5, -1, -1.5, -2, -3, -4, -6],
        "fp4_p_vals": [0, 0.5, 1, 1.5, 2, 3, 4, 6],
        "fp4_thresholds": [0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0],
        "thr_pos": [0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 7.0],
    }[_name]
    setattr(ref_qu, _name, _L([torch.tensor(_vals, dtype=torch.float32)]))


# ---------------------------------------------------------------------------------------------
def bits(t):
    """torch tensor -> numpy array of raw bits (exact round trip)."""
    if t is None:
        return np.zeros((0,), np.uint8)
    t = t.detach().contiguous()
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.view(torch.uint16).numpy().copy() if hasattr(torch, "uint16") else t.view(torch.int16).numpy().copy()
    if t.dtype in (torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e8m0fnu):
        return t.view(torch.uint8).numpy().copy()
    return t.numpy().copy()


MANIFEST = {}


def save(name, arrays, info):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    MANIFEST[name] = info
    print("saved", name, info.get("kernels_ok"), flush=True)


def run_kernels(layer, x, kernels):
    out, ok, failed = {}, [], {}
    for mt in kernels:
        if x.shape[0] > 1 and "GEMV" in mt:
            continue
        try:
            y = forward_functional._init_fn(x, layer.bias, layer.get_tensor_args(), layer.get_meta_args(),
                                            GEMLITE_MATMUL_TYPES_MAPPING[mt])
            out[f"y_{mt}_M{x.shape[0]}"] = bits(y)
            ok.append(f"{mt}:M{x.shape[0]}")
        except Exception as e:  # interpreter limitation -> recorded, not hidden
            failed[f"{mt}:M{x.shape[0]}"] = repr(e)[:160]
    return out, ok, failed


ALL = ["GEMV_REVSPLITK", "GEMV", "GEMV_SPLITK", "GEMM_SPLITK", "GEMM"]


def layer_case(name, layer, xs, kernels, extra_in):
    arrays = dict(extra_in)
    arrays["packed_W_q"] = bits(layer.W_q.data)
    arrays["packed_W_q_strides"] = np.array(layer.W_q.stride(), np.int64)
    arrays["packed_scales"] = bits(layer.scales.data)
    arrays["packed_scales_strides"] = np.array(layer.scales.stride() if layer.scales.dim() == 2 else (0, 0), np.int64)
    arrays["packed_zeros"] = bits(layer.zeros.data)
    arrays["meta_args"] = np.array(layer.get_meta_args(), np.int64)
    arrays["metadata_param"] = layer.metadata.data.numpy().copy()
    ok_all, failed_all = [], {}
    for i, x in enumerate(xs):
        arrays[f"x{i}"] = bits(x)
        out, ok, failed = run_kernels(layer, x, kernels)
        arrays.update(out)
        ok_all += ok
        failed_all.update(failed)
    info = {
        "kernels_ok": ok_all, "kernels_failed": failed_all,
        "scales_dtype": str(layer.scales.dtype), "zeros_dtype": str(layer.zeros.dtype),
        "W_q_dtype": str(layer.W_q.dtype), "W_q_shape": list(layer.W_q.shape),
        "scales_shape": list(layer.scales.shape), "zeros_shape": list(layer.zeros.shape),
        "x_dtype": str(xs[0].dtype),
    }
    save(name, arrays, info)


def gen_int_data(N, K, nbits, gs, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    W_q = torch.randint(0, 2 ** nbits, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    ng = N * K // gs
    scales = (torch.rand(ng, 1, generator=g) * 0.01 + 0.001).to(dtype)
    zeros = (torch.rand(ng, 1, generator=g) * (2 ** nbits - 1)).to(dtype)
    return W_q, scales, zeros


def xgen(M, K, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(M, K, generator=g) / 10.0).to(dtype)


# ---------------------------------------------------------------------------------------------
def case_bitpack():
    arrays, ok = {}, []
    g = torch.Generator().manual_seed(7)
    for nbits in (8, 4, 2, 1):
        W_q = torch.randint(0, 2 ** nbits, (16, 128), generator=g, dtype=torch.int32).to(torch.uint8)
        arrays[f"in_b{nbits}"] = W_q.numpy().copy()
        for pw in (8, 16, 32, 64):
            if pw < nbits:
                continue
            for tr in (True, False):
                o, eps = ref_bitpack.pack_weights_over_cols_torch(W_q, nbits, pw, tr)
                arrays[f"torch_b{nbits}_w{pw}_t{int(tr)}"] = o.contiguous().numpy().copy()
                try:
                    o2, eps2 = ref_bitpack.pack_weights_over_cols_triton(W_q, nbits, pw, tr)
                    arrays[f"triton_b{nbits}_w{pw}_t{int(tr)}"] = o2.contiguous().numpy().copy()
                    ok.append(f"triton_b{nbits}_w{pw}")
                except Exception as e:
                    MANIFEST.setdefault("_notes", {})[f"pack_triton_b{nbits}_w{pw}"] = repr(e)[:120]
    save("bitpack", arrays, {"kernels_ok": sorted(set(ok))})


def case_a16wn(name, nbits, gs, dtype, fma_mode, N=64, K=256, Ms=(1, 4), bias=False, seed=0, packing_bitwidth=None):
    W_q, scales, zeros = gen_int_data(N, K, nbits, gs, dtype, seed)
    gd = gemlite.core.TORCH_TO_DTYPE[dtype]
    layer = GemLiteLinearTriton(nbits, gs, K, N, gd, gd)
    b = None
    if bias:
        b = (torch.randn(N, generator=torch.Generator().manual_seed(seed + 5)) / 10).to(dtype)
    layer.pack(W_q, scales, zeros, b, fma_mode=fma_mode, packing_bitwidth=packing_bitwidth)
    xs = [xgen(M, K, dtype, 100 + M) for M in Ms]
    extra = {"in_W_q": W_q.numpy().copy(), "in_scales": bits(scales), "in_zeros": bits(zeros)}
    if b is not None:
        extra["in_bias"] = bits(b)
    layer_case(name, layer, xs, ALL, extra)


def case_a16w8_int8(name, post_scale, dtype=torch.float16, N=64, K=256):
    from gemlite.helper import A16W8_INT8
    g = torch.Generator().manual_seed(11)
    W = (torch.randn(N, K, generator=g) / 10).to(dtype)
    layer = A16W8_INT8(device="cpu", dtype=dtype, post_scale=post_scale).from_weights(W.clone())
    xs = [xgen(M, K, dtype, 200 + M) for M in (1, 4)]
    layer_case(name, layer, xs, ALL, {"in_W": bits(W)})


def case_int8_wn_scalar_zero(name, N=64, K=256, nbits=4):
    g = torch.Generator().manual_seed(21)
    W_q = torch.randint(0, 2 ** nbits - 1, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    sc = (torch.randn(N, 1, generator=g) * 1e-4).to(torch.float16)
    layer = GemLiteLinearTriton(nbits, K, K, N, DType.INT8, DType.FP32, scaled_activations=False)
    layer.pack(W_q, scales=sc, zeros=7, bias=None)
    xs = [torch.randint(-10, 10, (M, K), generator=g, dtype=torch.int32).to(torch.int8) for M in (1, 4)]
    layer_case(name, layer, xs, ALL, {"in_W_q": W_q.numpy().copy(), "in_scales": bits(sc)})


def case_a8w8_dynamic(name, fp8, N=64, K=256, dtype=torch.float16):
    from gemlite.helper import A8W8_INT8_dynamic, A8W8_FP8_dynamic
    g = torch.Generator().manual_seed(31)
    W = (torch.randn(N, K, generator=g) / 10).to(dtype)
    proc = A8W8_FP8_dynamic(device="cpu", dtype=dtype) if fp8 else A8W8_INT8_dynamic(device="cpu", dtype=dtype)
    layer = proc.from_weights(W.clone())
    xs = [xgen(M, K, dtype, 300 + M) for M in (1, 4)]
    arrays = {"in_W": bits(W)}
    # activation quantiser on its own (reference Triton kernel under the interpreter)
    try:
        for i, x in enumerate(xs):
            q, s = ref_qu.scale_activations_per_token_triton(x, torch.float8_e4m3fn if fp8 else torch.int8)
            arrays[f"actq_q{i}"] = bits(q)
            arrays[f"actq_s{i}"] = bits(s)
    except Exception as e:
        MANIFEST.setdefault("_notes", {})[name + "_actq"] = repr(e)[:160]
    layer_case(name, layer, xs, ALL, arrays)


def case_mxfp(name, nbits, N=64, K=256, dtype=torch.bfloat16):
    from gemlite.helper import A16W4_MXFP, A16W8_MXFP
    g = torch.Generator().manual_seed(41)
    W = (torch.randn(N, K, generator=g) / 10).to(dtype)
    lin = torch.nn.Linear(K, N, bias=False, dtype=dtype)
    lin.weight.data = W.clone()
    lin.weight.requires_grad = False
    proc = (A16W4_MXFP if nbits == 4 else A16W8_MXFP)(device="cpu", dtype=dtype)
    # quantiser is torch.compile'd in the reference; run its python body un-compiled on CPU
    q = ref_qu.WeightQuantizerMXFP(compute_dtype=dtype, device="cpu")
    fn = (q.quantize_mxfp4 if nbits == 4 else q.quantize_mxfp8)
    fn = getattr(fn, "_torchdynamo_orig_callable", None) or getattr(fn, "__wrapped__", None) or fn
    try:
        with _dynamo.config.patch(disable=True):
            W_qq, sc = (q.quantize_mxfp4(W, index=True) if nbits == 4 else q.quantize_mxfp8(W, index=True))
    except Exception:
        W_qq, sc = fn(q, W, index=True)
    W_qq, sc = W_qq.view(N, K), sc.view(N, K // 32)
    layer = proc.from_weights(W_qq, sc)
    xs = [xgen(M, K, dtype, 400 + M) for M in (1, 4, 16)]
    layer_case(name, layer, xs, ["GEMM_SPLITK", "GEMM", "GEMV"],
               {"in_W": bits(W), "q_codes": bits(W_qq), "q_scales": bits(sc)})


def _edge_rows(x):
    """Plant the cases the quantisers branch on into the first rows of x (in place): an all-zero group, a group whose
    maximum is an exact power of two times 6 / 448 (scale mantissa == 0 -> exponent not bumped) with elements sitting
    exactly on the e2m1 rounding thresholds (ties), a group of tiny values (exponent clamp at 2^-30) and a large one."""
    K = x.shape[1]
    x[0, 0:32] = 0
    if K >= 64:
        t = torch.tensor([6, 0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, -0.25, -0.75, -1.25, -1.75, -2.5, -3.5, -5.0, -6,
                          0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 0.26, 0.74, 5.01, 4.99, -0.0, 0.1, 0.2, 0.3, 0.4, 0.6]) * 0.125
        x[0, 32:64] = t.to(x.dtype)
    if K >= 96:
        x[0, 64:96] = (torch.arange(32) - 16).to(x.dtype) * 1e-7 if x.dtype != torch.float16 else (torch.arange(32) - 16).to(x.dtype) * 6e-8
    if K >= 128:
        x[0, 96:128] = (torch.arange(32) - 15.5).to(x.dtype) * (3000.0 if x.dtype == torch.float16 else 1e6)
    if K >= 160:
        x[0, 128:160] = torch.tensor([448.0] + [float(v) for v in range(1, 32)]).to(x.dtype) * 0.5
    return x


def case_actq_mx(name="actq_mx"):
    """MX / NV activation quantisers (quant_utils.py:503-954, the *_triton_v2 functions forward_functional calls)."""
    arrays, ok, failed = {}, [], {}
    fns = {"mxfp8": lambda t: ref_qu.scale_activations_mxfp8_triton_v2(t, torch.float8_e4m3fn),
           "mxfp4": ref_qu.scale_activations_mxfp4_triton_v2,
           "nvfp4": ref_qu.scale_activations_nvfp4_triton_v2}
    for dt, tag in ((torch.float16, "f16"), (torch.bfloat16, "bf16")):
        for M in (1, 5, 33):
            x = _edge_rows(xgen(M, 256, dt, 500 + M))
            arrays[f"x_{tag}_M{M}"] = bits(x)
            for kind, fn in fns.items():
                key = f"{kind}_{tag}_M{M}"
                try:
                    q, sc = fn(x.clone())
                    arrays["q_" + key] = bits(q)
                    arrays["s_" + key] = bits(sc)
                    ok.append(key)
                except Exception as e:
                    failed[key] = repr(e)[:200]
    save(name, arrays, {"kernels_ok": ok, "kernels_failed": failed})


def case_weight_quant_mx(name="weight_quant_mx"):
    """WeightQuantizerMXFP (quant_utils.py:70-225) incl. the window_size scale search: the reference's torch code on CPU."""
    arrays, ok = {}, []
    q = ref_qu.WeightQuantizerMXFP(compute_dtype=torch.bfloat16, device="cpu")
    for dt, tag in ((torch.bfloat16, "bf16"), (torch.float16, "f16")):
        g = torch.Generator().manual_seed(71)
        W = (torch.randn(48, 128, generator=g) / 10).to(dt)
        W[0, :32] = 0
        W[1, :32] = (torch.tensor([6, 0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0] * 4) * 0.0625).to(dt)
        arrays[f"W_{tag}"] = bits(W)
        with _dynamo.config.patch(disable=True):
            for ws in (0, 1, 2):
                for fn in ("quantize_mxfp4", "quantize_nvfp4"):
                    for index in (True, False):
                        Wq, sc = getattr(q, fn)(W.clone(), window_size=ws, index=index)
                        key = f"{fn}_{tag}_w{ws}_i{int(index)}"
                        arrays["q_" + key] = bits(Wq)
                        arrays["s_" + key] = bits(sc)
                        ok.append(key)
            for index in (True, False):
                Wq, sc = q.quantize_mxfp8(W.clone(), index=index)
                key = f"quantize_mxfp8_{tag}_i{int(index)}"
                arrays["q_" + key] = bits(Wq)
                arrays["s_" + key] = bits(sc)
                ok.append(key)
            Wq, sc = q.quantize_mxfp4(W.clone(), window_size=1, index=True)
            arrays[f"deq_mxfp4_{tag}"] = bits(q.dequantize(Wq, sc, shape=W.shape, dtype=dt))
    save(name, arrays, {"kernels_ok": ok})


def case_mx_dynamic(name, proc_name, kwargs, quant, N=64, K=256, dtype=torch.bfloat16):
    """MXFP / NVFP weights with dynamically quantised activations (helper.py:660-945)."""
    from gemlite import helper as ref_helper
    g = torch.Generator().manual_seed(51)
    W = (torch.randn(N, K, generator=g) / 10).to(dtype)
    q = ref_qu.WeightQuantizerMXFP(compute_dtype=dtype, device="cpu")
    with _dynamo.config.patch(disable=True):
        W_qq, sc = getattr(q, quant)(W, index=True)
    gs = 16 if quant == "quantize_nvfp4" else 32
    W_qq, sc = W_qq.view(N, K), sc.view(N, K // gs)
    layer = getattr(ref_helper, proc_name)(device="cpu", dtype=dtype, **kwargs).from_weights(W_qq, scales=sc)
    xs = [_edge_rows(xgen(M, K, dtype, 600 + M)) for M in (1, 4, 33)]
    layer_case(name, layer, xs, ["GEMM_SPLITK", "GEMM"],
               {"in_W": bits(W), "q_codes": bits(W_qq), "q_scales": bits(sc)})


def main():
    jobs = [
        ("bitpack", case_bitpack, ()),
        ("a16w4_gs64_fp16_fma", case_a16wn, ("a16w4_gs64_fp16_fma", 4, 64, torch.float16, True)),
        ("a16w4_gs128_fp16_fma_bias", case_a16wn, ("a16w4_gs128_fp16_fma_bias", 4, 128, torch.float16, True, 128, 512, (1, 4, 16), True, 1)),
        ("a16w4_gs64_fp16_nofma", case_a16wn, ("a16w4_gs64_fp16_nofma", 4, 64, torch.float16, False, 64, 256, (1, 4), False, 2)),
        ("a16w4_gs64_bf16_fma", case_a16wn, ("a16w4_gs64_bf16_fma", 4, 64, torch.bfloat16, True, 64, 256, (1, 4), False, 3)),
        ("a16w2_gs64_fp16_fma", case_a16wn, ("a16w2_gs64_fp16_fma", 2, 64, torch.float16, True, 64, 256, (1, 4), False, 4)),
        ("a16w1_gs64_fp16_fma", case_a16wn, ("a16w1_gs64_fp16_fma", 1, 64, torch.float16, True, 64, 256, (1, 4), False, 5)),
        ("a16w8_gs64_fp16_fma", case_a16wn, ("a16w8_gs64_fp16_fma", 8, 64, torch.float16, True, 64, 256, (1, 4), False, 6)),
        ("a16w4_gs64_fp16_fma_pw8", case_a16wn, ("a16w4_gs64_fp16_fma_pw8", 4, 64, torch.float16, True, 64, 256, (1, 4), False, 8, 8)),
        ("a16w4_channelwise_fp16", case_a16wn, ("a16w4_channelwise_fp16", 4, 256, torch.float16, True, 64, 256, (1, 4), False, 9)),
        ("a16w8_int8_prescale", case_a16w8_int8, ("a16w8_int8_prescale", False)),
        ("a16w8_int8_postscale", case_a16w8_int8, ("a16w8_int8_postscale", True)),
        ("int8_w4_scalar_zero", case_int8_wn_scalar_zero, ("int8_w4_scalar_zero",)),
        ("a8w8_int8_dynamic", case_a8w8_dynamic, ("a8w8_int8_dynamic", False)),
        ("a8w8_fp8_dynamic", case_a8w8_dynamic, ("a8w8_fp8_dynamic", True)),
        ("a16w4_mxfp_bf16", case_mxfp, ("a16w4_mxfp_bf16", 4)),
        ("a16w8_mxfp_bf16", case_mxfp, ("a16w8_mxfp_bf16", 8)),
        ("actq_mx", case_actq_mx, ()),
        ("weight_quant_mx", case_weight_quant_mx, ()),
        ("a8w8_mxfp_dynamic_post", case_mx_dynamic, ("a8w8_mxfp_dynamic_post", "A8W8_MXFP_dynamic", {"post_scale": True}, "quantize_mxfp8")),
        ("a8w8_mxfp_dynamic", case_mx_dynamic, ("a8w8_mxfp_dynamic", "A8W8_MXFP_dynamic", {"post_scale": False}, "quantize_mxfp8")),
        ("a8w4_mxfp_dynamic", case_mx_dynamic, ("a8w4_mxfp_dynamic", "A8W4_MXFP_dynamic", {"post_scale": False}, "quantize_mxfp4")),
        ("a8w4_mxfp_dynamic_post", case_mx_dynamic, ("a8w4_mxfp_dynamic_post", "A8W4_MXFP_dynamic", {"post_scale": True}, "quantize_mxfp4")),
        ("a4w4_mxfp_dynamic", case_mx_dynamic, ("a4w4_mxfp_dynamic", "A4W4_MXFP_dynamic", {}, "quantize_mxfp4")),
        ("a4w4_nvfp_dynamic", case_mx_dynamic, ("a4w4_nvfp_dynamic", "A4W4_NVFP_dynamic", {}, "quantize_nvfp4")),
    ]
    only = set(sys.argv[1:])
    for name, fn, args in jobs:
        if only and name not in only:
            continue
        try:
            fn(*args)
        except Exception as e:
            traceback.print_exc()
            MANIFEST.setdefault("_failed_cases", {})[name] = repr(e)[:200]
    mpath = os.path.join(HERE, "MANIFEST.json")
    old = {}
    if only and os.path.exists(mpath):
        old = json.load(open(mpath))
    old.update(MANIFEST)
    old["_reference"] = {"version": gemlite.__version__, "triton": triton.__version__, "torch": torch.__version__,
                         "mode": "TRITON_INTERPRET=1 (CPU), stub driver, unmodified reference sources; interpreter float "
                                 "down-casts replaced by RNE / satfinite, numpy stand-ins for libdevice.round and tl.dot_scaled"}
    json.dump(old, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
