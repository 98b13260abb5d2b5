"""CPU: the C-ABI library builds, loads and exports exactly what include/gemlite_b200.h declares;
host-side logic (validation, mode resolution inputs, dispatch table) behaves like the reference."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import gemlite_b200
from gemlite_b200 import DType, GemLiteLinear, _cabi, get_matmul_type
from gemlite_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    lib_path = build()
    assert os.path.exists(lib_path)
    header = open(os.path.join(ROOT, "include", "gemlite_b200.h")).read()
    declared = sorted(set(re.findall(r"GLB_API\s+[\w\s\*]+?\b(glb_\w+)\s*\(", header)))
    assert declared == sorted(_cabi.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared, (exported, declared)
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.glb_version() == 1


def test_struct_layout_matches_header():
    src = '#include "%s/include/gemlite_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%%zu %%zu %%zu %%zu\\n", sizeof(glb_linear), offsetof(glb_linear, scales), offsetof(glb_linear, K), offsetof(glb_linear, W_group_mode));}' % ROOT
    exe = "/tmp/_glb_sz"
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    size, o_scales, o_k, o_wgm = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    assert size == ctypes.sizeof(_cabi.GlbLinear)
    assert o_scales == _cabi.GlbLinear.scales.offset
    assert o_k == _cabi.GlbLinear.K.offset
    assert o_wgm == _cabi.GlbLinear.W_group_mode.offset


def test_dtype_enum_wire_values():
    want = dict(FP32=0, FP16=1, BF16=2, FP8=3, INT8=4, UINT8=5, INT32=6, UINT32=7, FP8e5=8, INT16=9, UINT16=10,
                INT64=11, FP8e4nuz=12, FP8e5nuz=13, MXFP16=14, MXBF16=15, MXFP8=16, MXFP4=17, NVFP4=18, E8M0=19)
    for k, v in want.items():
        assert DType[k].value == v


def test_dispatch_table_matches_reference():
    # core.py:100-114
    assert get_matmul_type(1, 4) == "GEMV_REVSPLITK"
    assert get_matmul_type(1, 8) == "GEMV_SPLITK"
    assert get_matmul_type(1, 4, True) == "GEMM_SPLITK"
    assert get_matmul_type(2, 4) == "GEMM_SPLITK" and get_matmul_type(64, 4) == "GEMM_SPLITK"
    assert get_matmul_type(65, 4) == "GEMM"
    assert gemlite_b200.GEMLITE_MATMUL_TYPES == ["GEMV", "GEMV_REVSPLITK", "GEMV_SPLITK", "GEMM_SPLITK", "GEMM"]


def test_constructor_validation_like_reference():
    with pytest.raises(NotImplementedError):
        GemLiteLinear(W_nbits=3, group_size=64, in_features=256, out_features=64)
    with pytest.raises(NotImplementedError):
        GemLiteLinear(W_nbits=4, group_size=64, in_features=250, out_features=64)
    with pytest.raises(NotImplementedError):
        GemLiteLinear(W_nbits=4, group_size=8, in_features=256, out_features=64)
    with pytest.raises(NotImplementedError):
        GemLiteLinear(W_nbits=4, group_size=48, in_features=256, out_features=64)
    layer = GemLiteLinear(W_nbits=4, group_size=None, in_features=256, out_features=64)
    assert layer.group_size == 1 and layer.unpack_mask == 15
    # fp16 inputs never get dynamic activation scaling (core.py:293-296)
    assert GemLiteLinear(4, 64, 256, 64, DType.FP16, DType.FP16, scaled_activations=True).scaled_activations is False
    assert GemLiteLinear(8, 256, 256, 64, DType.INT8, DType.FP16, scaled_activations=True).scaled_activations is True
    assert GemLiteLinear(8, 256, 256, 64, DType.INT8, DType.FP16).acc_dtype == DType.INT32


def test_no_cpu_fallback():
    """The product must fail loudly off-GPU instead of computing on the host."""
    layer = GemLiteLinear(4, 64, 256, 64, DType.FP16, DType.FP16)
    W_q = torch.randint(0, 16, (64, 256), dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer.pack(W_q, torch.ones(256, 1, dtype=torch.float16), torch.ones(256, 1, dtype=torch.float16))


def test_product_does_not_import_oracle():
    import ast
    pkg = os.path.join(ROOT, "gemlite_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, fn)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), fn


def test_matmul_rejects_bad_descriptor_without_gpu():
    lib = _cabi.load()
    d = _cabi.GlbLinear()
    rc = lib.glb_matmul(ctypes.byref(d), None, None, None, 1, -1, None, 0, None, None)
    assert rc == -1 and b"no weights" in lib.glb_last_error()
    assert lib.glb_pack_cols(None, 4, 64, 3, 32, None, 1, None) == -1
    assert b"nbits" in lib.glb_last_error()


def _a16w4_desc(N, K, gs=128):
    """A descriptor shaped like a packed A16W4 layer (fake, aligned device addresses: nothing is dereferenced)."""
    d = _cabi.GlbLinear()
    d.W_q = 0x7f0000000000; d.wq_rows, d.wq_cols = K // 8, N; d.wq_stride0, d.wq_stride1 = N, 1
    d.wq_dtype = DType.INT32.value
    d.scales = 0x7f0100000000; d.scales_numel = (K // gs) * N; d.s_stride0, d.s_stride1 = N, 1
    d.scales_dtype = DType.FP16.value
    d.zeros = 0x7f0200000000; d.zeros_numel = (K // gs) * N; d.z_stride0, d.z_stride1 = N, 1
    d.zeros_dtype = DType.FP16.value
    d.K, d.N = K, N
    (d.scaled_activations, d.W_nbits, d.group_size, d.unpack_mask, d.elements_per_sample, d.input_dtype, d.output_dtype,
     d.acc_dtype, d.meta_dtype, d.channel_scale_mode, d.W_group_mode, d.data_contiguous) = (
        0, 4, gs, 15, 8, DType.FP16.value, DType.FP16.value, DType.FP32.value, DType.FP16.value, 0, 4, 1)
    return d


def test_workspace_plan_without_gpu():
    """glb_workspace_bytes is pure host logic: zero for the decode kernels and for full-size prefill, counters + fp32
    partial tiles for the split-K range 16 < M <= 128 of a narrow layer (tiles * splits <= SM pairs)."""
    lib = _cabi.load()
    d = _a16w4_desc(4096, 4096)
    assert lib.glb_workspace_bytes(ctypes.byref(d), 1, -1) == 0          # decode kernel
    assert lib.glb_workspace_bytes(ctypes.byref(d), 16, -1) == 0
    assert lib.glb_workspace_bytes(ctypes.byref(d), 4096, -1) == 0       # 256 tiles: no split
    assert lib.glb_workspace_bytes(ctypes.byref(d), 256, -1) == 0        # split-K is not used above M = 128
    # 16 < M <= 128 goes to the batched-decode kernel (gl_gemm_mid.cu, no scratch) unless "GEMM" is requested by name,
    # which keeps the CTA-pair tcgen05 kernel and its split-K plan
    assert lib.glb_workspace_bytes(ctypes.byref(d), 32, -1) == 0 and lib.glb_workspace_bytes(ctypes.byref(d), 128, -1) == 0
    GEMM = 4
    w32 = lib.glb_workspace_bytes(ctypes.byref(d), 32, GEMM)
    # fixed 64 KB counter region (the same for every layer: one workspace serves all layers of a stream, so the
    # partial tiles of one layer must never overlap the words another layer reads as counters)
    # + 16 tiles x 4 splits x 2 CTAs x 128 rows x 64 columns fp32
    assert w32 == 65536 + 16 * 4 * 2 * 128 * 64 * 4, w32
    w128 = lib.glb_workspace_bytes(ctypes.byref(d), 128, GEMM)
    assert w128 == 65536 + 16 * 4 * 2 * 128 * 128 * 4, w128
    # a layer with another tile count gets the same counter region
    other = _a16w4_desc(2048, 4096)
    assert lib.glb_workspace_bytes(ctypes.byref(other), 32, GEMM) == 65536 + 8 * 8 * 2 * 128 * 64 * 4
    wide = _a16w4_desc(16384, 4096)
    assert lib.glb_workspace_bytes(ctypes.byref(wide), 64, GEMM) == 0    # 64 tiles already fill the 74 SM pairs


def _mx_desc(N, K, input_dtype, nbits, gs, csm, scales_dtype):
    """A descriptor shaped like an MX layer with dynamically quantised activations (fake addresses, never dereferenced)."""
    d = _cabi.GlbLinear()
    eps = 2 if nbits == 4 else 1
    d.W_q = 0x7f0000000000; d.wq_rows, d.wq_cols = K // eps, N; d.wq_stride0, d.wq_stride1 = 1, K // eps
    d.wq_dtype = (DType.UINT8 if nbits == 4 else DType.FP8).value
    d.scales = 0x7f0100000000; d.scales_numel = (K // gs) * N; d.s_stride0, d.s_stride1 = 1, N
    d.scales_dtype = scales_dtype
    d.zeros = None; d.zeros_numel = 0
    d.zeros_dtype = DType.INT32.value
    d.K, d.N = K, N
    (d.scaled_activations, d.W_nbits, d.group_size, d.unpack_mask, d.elements_per_sample, d.input_dtype, d.output_dtype,
     d.acc_dtype, d.meta_dtype, d.channel_scale_mode, d.W_group_mode, d.data_contiguous) = (
        1, nbits, gs, 2 ** nbits - 1, eps, input_dtype, DType.BF16.value, DType.FP32.value, scales_dtype, csm, 0, 0)
    return d


def test_mx_dynamic_descriptor_validation_without_gpu():
    """glb_matmul validates before it touches the device: which (input dtype, channel_scale_mode, group size) an MX
    layer may carry (helper.py:660-945) and that modes 2 / 4 come with activation scales."""
    lib = _cabi.load()
    fake = 0x7f0300000000
    u8 = DType.UINT8.value

    def call(d, scales_x):
        return lib.glb_matmul(ctypes.byref(d), fake, scales_x, fake, 1, -1, None, 0, None, None), lib.glb_last_error()

    # micro-scaled activations without their scales: invalid argument (-1)
    rc, msg = call(_mx_desc(256, 512, DType.MXFP8.value, 8, 32, 4, u8), None)
    assert rc == -1 and b"activation scales" in msg, (rc, msg)
    rc, msg = call(_mx_desc(256, 512, DType.MXFP8.value, 4, 32, 2, u8), None)
    assert rc == -1 and b"activation scales" in msg, (rc, msg)
    # combinations the reference cannot produce: unsupported (-2)
    for d in (_mx_desc(256, 512, DType.MXFP4.value, 4, 32, 2, u8),          # per-token post-scale exists for MXFP8 only
              _mx_desc(256, 512, DType.MXBF16.value, 4, 32, 4, u8),         # weight-only MX has no activation scales
              _mx_desc(256, 512, DType.MXFP4.value, 4, 16, 4, u8),          # MXFP4 groups are 32 wide
              _mx_desc(256, 512, DType.NVFP4.value, 4, 32, 4, DType.FP8.value)):   # NVFP4 groups are 16 wide
        rc, msg = call(d, fake)
        assert rc == -2, (rc, msg)


def test_workspace_plan_of_the_round2_routes_without_gpu():
    """Host-side routing of the kernels added in round 2 (glb_workspace_bytes mirrors glb_matmul's dispatch): the
    block-scaled MX kernel, the packed-width decode kernel and the A16W8 batched-decode kernel need no scratch; the
    expansion path (A8W4, or A4W4 MXFP4 at M <= 4) needs room for the expanded activations behind the 64 KB counter region;
    packed 2-bit layers above the decode range take the pair kernel's split-K plan."""
    lib = _cabi.load()
    u8 = DType.UINT8.value
    ws = lambda d, M, mt=-1: lib.glb_workspace_bytes(ctypes.byref(d), M, mt)   # noqa: E731
    # A8W8 MXFP8 and A4W4 NVFP4: block-scaled kernel at every M
    for d in (_mx_desc(4096, 4096, DType.MXFP8.value, 8, 32, 4, u8), _mx_desc(4096, 4096, DType.NVFP4.value, 4, 16, 4, DType.FP8.value)):
        assert [ws(d, M) for M in (1, 64, 4096)] == [0, 0, 0]
    # A4W4 MXFP4 (bf16): streaming kernel behind the bf16 expansion at M <= 4, block-scaled above
    a4 = _mx_desc(4096, 4096, DType.MXFP4.value, 4, 32, 4, u8)
    # (the expanded activations sit behind the inner launch's own scratch, never inside the first 64 KB: the split-K counters)
    for M in (1, 4):
        assert ws(a4, M) >= 65536 + M * 4096 * 2 and (ws(a4, M) - M * 4096 * 2) % 256 == 0, (M, ws(a4, M))
    assert ws(a4, 5) == 0 and ws(a4, 4096) == 0
    # a width the block-scaled kernel does not take (N % 256 != 0) stays on the expansion path
    odd = _mx_desc(4096 + 128, 4096, DType.MXFP4.value, 4, 32, 4, u8)
    assert ws(odd, 64) == 0 or ws(odd, 64) >= 65536
    # packed 2-bit grouped layer: decode kernel (no scratch) to M = 8, pair kernel + split-K plan above
    w2 = _a16w4_desc(4096, 4096, gs=64)
    w2.W_nbits, w2.elements_per_sample, w2.unpack_mask, w2.wq_rows = 2, 16, 3, 4096 // 16
    assert ws(w2, 1) == 0 and ws(w2, 8) == 0
    assert ws(w2, 16) == 65536 + 16 * 4 * 2 * 128 * 64 * 4        # same plan as the 4-bit layer by name (tests above)
    assert ws(w2, 4096) == 0
