"""Parity at BASELINE.json's FULL sizes against the CPU oracle (no reference package needed: this file must stay green on
a box without baseline/_ref; tests/test_gpu_vs_reference.py adds the live reference on the same cases).

  C3  A16W4 gs=128, M=4096, K=N=8192, fp16 and bf16 (64 k-stages per tile, 16+ ring wraps of the pair kernel)
  C4  A8W8 int8 / fp8 dynamic, 4096^2, M in {1, 16, 64, 128, 129, 1024, 4096}: int8 bit-exact, fp8 <= 2e-3
  C5  A16W4 MXFP4 gs=32, K=N=16384, M in {1, 4, 16, 4096}
The oracle computes sampled token rows (all output columns) -- a full 4096 x 8192 x 8192 fp32 product on the host would
take minutes; the rows are drawn with a fixed seed and always include the first and last row of every 128-token block
boundary the kernels tile on."""
import pytest
import torch

import gemlite_b200
import gemlite_b200.helper  # noqa: F401
from gemlite_b200 import _cabi
from oracle import gemlite_oracle as O
import refdual
from test_gpu_parity import REL, oracle_from_layer, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rows(M, n):
    if M <= n:
        return torch.arange(M)
    g = torch.Generator().manual_seed(M)
    fixed = torch.tensor([0, 127, 128, 255, 256, M // 2 - 1, M // 2, M - 129, M - 128, M - 1])
    rnd = torch.randperm(M, generator=g)[: n - fixed.numel()]
    return torch.unique(torch.cat([fixed.clamp(0, M - 1), rnd]))


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_c3_prefill_full_size(dtype):
    b = dict(kind="pack_int", nbits=4, gs=128, N=8192, K=8192, dtype=dtype, seed=31)
    layer = refdual.build_layer(gemlite_b200, b)
    p = oracle_from_layer(layer)
    case = dict(build=b)
    x = refdual.x_for(case, 4096, 0)
    y = layer(x)
    torch.cuda.synchronize()
    assert _cabi.last_kernel() == "gemm_a16w4_tcgen05_2cta"
    rows = _rows(4096, 40)
    yo = O.forward(p, x[rows.to(DEV)].cpu())
    e = relerr(y[rows.to(DEV)], yo)
    assert e <= REL[refdual.TD[dtype]], (dtype, e)
    # size-independent property on the whole output: linearity in x (power-of-two scaling is exact in fp16 / bf16)
    y2 = layer(x * 0.5)
    assert torch.equal(y2.float() * 2, y.float()) or relerr(y2.float() * 2, y.float()) < 1e-3
    assert torch.isfinite(y.float()).all()


@pytest.mark.parametrize("kind", ["int8", "fp8"])
def test_c4_a8w8_full_size(kind):
    cls = "A8W8_INT8_dynamic" if kind == "int8" else "A8W8_FP8_dynamic"
    b = dict(kind="helper_w", cls=cls, N=4096, K=4096, dtype="float16", seed=41)
    layer = refdual.build_layer(gemlite_b200, b)
    p = oracle_from_layer(layer)
    case = dict(build=b)
    for i, M in enumerate([1, 16, 64, 128, 129, 1024, 4096]):
        x = refdual.x_for(case, M, i)
        y = layer(x)
        torch.cuda.synchronize()
        rows = _rows(M, 24)
        xr = x[rows.to(DEV)]
        if kind == "int8":
            # exactness is checked where it is defined: on identical quantised activations.  The product's own
            # quantiser output feeds the oracle's int32 GEMM + epilogue (cvt.rn.f32.s32(acc) * (s_x * s_w) -> fp16)
            q, s = gemlite_b200.quant_utils.scale_activations_per_token(xr, w_dtype=torch.int8)
            W = layer.W_q.data.t().contiguous().to(torch.int32)                      # (N, K) int8 codes
            acc = (q.to(torch.int32).cpu() @ W.cpu().t())                         # exact int32
            sw = layer.scales.data.view(1, -1).float().cpu()
            ref = (acc.float() * (s.float().cpu().view(-1, 1) * sw)).to(torch.float16)
            assert torch.equal(y[rows.to(DEV)].cpu(), ref), (kind, M, _cabi.last_kernel())
        else:
            yo = O.forward(p, xr.cpu())
            e = relerr(y[rows.to(DEV)], yo)
            assert e <= 2e-2, (kind, M, _cabi.last_kernel(), e)   # an fp8 activation code may round the other way
            q, s = gemlite_b200.quant_utils.scale_activations_per_token(xr, w_dtype=torch.float8_e4m3fn)
            W = layer.W_q.data.t().contiguous().float().cpu()
            ref = (q.float().cpu() @ W.t()) * (s.float().cpu().view(-1, 1) * layer.scales.data.view(1, -1).float().cpu())
            e2 = relerr(y[rows.to(DEV)], ref)
            assert e2 <= 2e-3, (kind, M, _cabi.last_kernel(), e2)


def test_c5_mxfp4_full_size():
    from gemlite_b200.quant_utils import WeightQuantizerMXFP
    N = K = 16384
    W = refdual.gen_w(N, K, torch.bfloat16, 51).to(DEV)
    W_q, sc = WeightQuantizerMXFP(compute_dtype=torch.bfloat16, device=DEV).quantize_mxfp4(W, index=True)
    del W
    layer = gemlite_b200.helper.A16W4_MXFP(device=DEV, dtype=torch.bfloat16).from_weights(W_q.view(N, K), sc.view(N, K // 32))
    p = oracle_from_layer(layer)
    case = dict(build=dict(K=K, dtype="bfloat16"))
    for i, M in enumerate([1, 4, 16, 4096]):
        x = refdual.x_for(case, M, i)
        y = layer(x)
        torch.cuda.synchronize()
        rows = _rows(M, 12)
        yo = O.forward(p, x[rows.to(DEV)].cpu())
        e = relerr(y[rows.to(DEV)], yo)
        assert e <= REL[torch.bfloat16], (M, _cabi.last_kernel(), e)
        assert torch.isfinite(y.float()).all()
