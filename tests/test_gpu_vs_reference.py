"""SURVEY 8(c) / VERDICT r1 row g: the CUDA path against the UNMODIFIED reference (its Triton kernels) on the same GPU,
on identical inputs, at BASELINE.json's full sizes.

The reference runs in a child process (tools/ref_runner.py, package under baseline/_ref); both sides execute
tests/refdual.py.  Bars:
  * packed tensors + meta_args: bit-equal;
  * INT8 x INT8: outputs `torch.equal` the reference's GEMM and GEMV_SPLITK outputs; per-token int8 / fp8 codes and
    scales `torch.equal` the reference's Triton quantiser;
  * floating point: every output within the stated tolerance of the fp32 CPU oracle (REL / REL_DECODE of
    test_gpu_parity.py) whether or not the reference ran; where it ran: mean|ours - ref| < 1e-3 (the reference's own test
    bar, test_gemlitelineartriton.py) -- unless the REFERENCE is the outlier against the oracle (its GEMVs accumulate in
    fp16 with atomics; its default GEMV_SPLITK / GEMV_REVSPLITK configs break on grouped 8-bit / 1-bit weights), which
    is recorded as `reference_outlier`; and our max / mean error against the oracle must not exceed the reference's
    (5 % + 1 output-ulp slack).  The factorised decode kernel is judged, with the reference, against the exact fp64
    product of the stored parameters instead (it skips the fp16 rounding of every dequantised weight that the oracle
    and the reference's GEMM perform, so it is closer to that product and further from the oracle).
A JSON report of every number goes to gpurun_out/vs_reference.json (copied to profiles/ by hand)."""
import json
import os
import subprocess
import sys
import tempfile

import pytest
import torch

import gemlite_b200
import gemlite_b200.helper  # noqa: F401
from gemlite_b200 import _cabi
from oracle import gemlite_oracle as O
import refdual
from test_gpu_parity import REL, REL_DECODE, oracle_from_layer

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "gemlite")
DEV = "cuda:0"
REPORT = {}

I4 = dict(kind="pack_int", nbits=4, gs=128)
CASES = [
    # ---- C2: A16W4 gs=128 4096^2, M=1, every kernel name + auto dispatch; fp16 and bf16
    dict(name="c2_fp16", build=dict(I4, N=4096, K=4096, dtype="float16", seed=0), Ms=[1],
         kernels=["GEMV_REVSPLITK", "GEMV", "GEMM_SPLITK", "GEMM", "AUTO"]),
    dict(name="c2_bf16", build=dict(I4, N=4096, K=4096, dtype="bfloat16", seed=1), Ms=[1],
         kernels=["GEMV_REVSPLITK", "GEMM_SPLITK", "AUTO"]),
    dict(name="c2_batched_fp16", build=dict(I4, N=4096, K=4096, dtype="float16", seed=2, bias=True),
         Ms=[2, 8, 16, 17, 64, 128, 300], kernels=["AUTO", "GEMM"], keep_packed=False),
    # ---- C3: prefill M=4096 K=N=8192
    dict(name="c3_fp16", build=dict(I4, N=8192, K=8192, dtype="float16", seed=3), Ms=[4096], kernels=["AUTO"],
         keep_packed=False, sample_rows=48),
    dict(name="c3_bf16", build=dict(I4, N=8192, K=8192, dtype="bfloat16", seed=4), Ms=[4096], kernels=["AUTO"],
         keep_packed=False, sample_rows=48),
    # ---- C4: A8W8 dynamic 4096^2, M sweep
    dict(name="c4_int8", build=dict(kind="helper_w", cls="A8W8_INT8_dynamic", N=4096, K=4096, dtype="float16", seed=5),
         Ms=[1, 16, 64, 128, 129, 1024, 4096], kernels=["AUTO", "GEMM", "GEMV_SPLITK"], actq=torch.int8,
         sample_rows=32, exact=("GEMM", "GEMV_SPLITK")),
    dict(name="c4_fp8", build=dict(kind="helper_w", cls="A8W8_FP8_dynamic", N=4096, K=4096, dtype="float16", seed=6),
         Ms=[1, 16, 64, 128, 129, 1024, 4096], kernels=["AUTO", "GEMM", "GEMV_SPLITK"], actq=torch.float8_e4m3fn,
         sample_rows=32),
    # ---- C5: MXFP4 gs=32 K=N=16384 (weights quantised once, shipped to both)
    dict(name="c5_mxfp4", build=dict(kind="mx_weights", cls="A16W4_MXFP", tensors="c5", N=16384, K=16384,
                                     dtype="bfloat16", seed=7), Ms=[1, 4, 16, 4096], kernels=["AUTO"],
         keep_packed=False, sample_rows=16),
    # ---- C1's GPU twin and the other weight formats of helper.py
    dict(name="a16w8_int8", build=dict(kind="helper_w", cls="A16W8_INT8", N=4096, K=4096, dtype="float16", seed=8),
         Ms=[1, 64, 512], kernels=["AUTO"]),
    dict(name="a16w8_int8_post", build=dict(kind="helper_w", cls="A16W8_INT8", N=1024, K=4096, dtype="float16", seed=9,
                                             kwargs=dict(post_scale=True)), Ms=[1, 64, 200], kernels=["AUTO"]),
    dict(name="a16w8_fp8", build=dict(kind="helper_w", cls="A16W8_FP8", N=4096, K=4096, dtype="bfloat16", seed=10),
         Ms=[1, 64, 512], kernels=["AUTO"]),
    dict(name="a16w2_hqq", build=dict(kind="hqq_int", cls="A16W2_HQQ_INT", nbits=2, gs=64, N=4096, K=4096,
                                      dtype="float16", seed=11), Ms=[1, 64, 512], kernels=["AUTO"]),
    dict(name="a16w1_hqq", build=dict(kind="hqq_int", cls="A16W1_HQQ_INT", nbits=1, gs=64, N=1024, K=4096,
                                      dtype="float16", seed=12), Ms=[1, 33, 200], kernels=["AUTO"]),
    dict(name="a16w8_hqq", build=dict(kind="hqq_int", cls="A16W8_HQQ_INT", nbits=8, gs=128, N=1024, K=4096,
                                      dtype="float16", seed=13), Ms=[1, 33, 200], kernels=["AUTO"]),
    dict(name="a16w4_channelwise", build=dict(kind="hqq_int", cls="A16W4_HQQ_INT", nbits=4, gs=4096, N=1024, K=4096,
                                              dtype="float16", seed=14), Ms=[1, 33, 200], kernels=["AUTO"]),
    dict(name="fp16xfp16", build=dict(kind="fp16", N=1024, K=4096, dtype="float16", seed=15), Ms=[1, 4, 200],
         kernels=["GEMV_REVSPLITK", "GEMV", "GEMV_SPLITK", "GEMM_SPLITK", "GEMM", "AUTO"]),
    dict(name="bitnet_a16w158", build=dict(kind="bitnet", cls="A16W158_INT", N=1024, K=4096, dtype="float16", seed=16),
         Ms=[1, 33, 200], kernels=["AUTO"]),
    # ---- dynamic MX / NV layers
    dict(name="a8w8_mxfp_post", build=dict(kind="mx_dynamic", cls="A8W8_MXFP_dynamic", N=2048, K=4096, dtype="bfloat16",
                                           seed=17, kwargs=dict(post_scale=True)), Ms=[1, 16, 200], kernels=["AUTO"]),
    dict(name="a8w8_mxfp", build=dict(kind="mx_dynamic", cls="A8W8_MXFP_dynamic", N=2048, K=4096, dtype="bfloat16",
                                      seed=18, kwargs=dict(post_scale=False)), Ms=[1, 16, 200], kernels=["AUTO"]),
    dict(name="a8w4_mxfp", build=dict(kind="mx_dynamic", cls="A8W4_MXFP_dynamic", N=2048, K=4096, dtype="bfloat16",
                                      seed=19, kwargs=dict(post_scale=False)), Ms=[1, 16, 200], kernels=["AUTO"]),
    dict(name="a4w4_mxfp", build=dict(kind="mx_dynamic", cls="A4W4_MXFP_dynamic", N=2048, K=4096, dtype="bfloat16",
                                      seed=20), Ms=[1, 16, 200], kernels=["AUTO"]),
    dict(name="a4w4_nvfp", build=dict(kind="mx_dynamic", cls="A4W4_NVFP_dynamic", N=2048, K=4096, dtype="bfloat16",
                                      seed=21), Ms=[1, 16, 200], kernels=["AUTO"]),
]
CASE_BY_NAME = {c["name"]: c for c in CASES}
if os.environ.get("GLB_VSREF_ONLY"):
    CASES = [c for c in CASES if any(k in c["name"] for k in os.environ["GLB_VSREF_ONLY"].split(","))]


def _c5_tensors():
    from gemlite_b200.quant_utils import WeightQuantizerMXFP
    W = refdual.gen_w(16384, 16384, torch.bfloat16, 7).to(DEV)
    q = WeightQuantizerMXFP(compute_dtype=torch.bfloat16, device=DEV)
    W_q, sc = q.quantize_mxfp4(W, index=True)
    return {"W_q": W_q.view(16384, 16384).cpu(), "scales": sc.view(16384, 512).cpu()}


@pytest.fixture(scope="module")
def both():
    """Runs every case through the reference (child process) once; ours is run lazily per test."""
    if not os.path.isdir(REF_DIR):
        pytest.skip("baseline/_ref/gemlite is absent (install: see DESIGN.md section 4)")
    tensors = {}
    if any(c["build"].get("tensors") == "c5" for c in CASES):
        tensors["c5"] = _c5_tensors()
        torch.cuda.empty_cache()
    tmp = tempfile.mkdtemp(prefix="vsref_")
    job, out = os.path.join(tmp, "job.pt"), os.path.join(tmp, "out.pt")
    torch.save({"cases": CASES, "tensors": tensors, "autotune": "default",
                "budget_s": float(os.environ.get("GLB_VSREF_BUDGET_S", "900"))}, job)
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    log = os.path.join(ROOT, "gpurun_out", "vs_reference_child.log")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    with open(log, "w") as lf:
        rc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_runner.py"), job, out], stdout=lf,
                            stderr=subprocess.STDOUT, env=env, cwd=tmp, timeout=1500).returncode
    assert os.path.exists(out), f"reference child produced nothing (rc={rc}); see {log}"
    ref = torch.load(out, weights_only=False)
    REPORT["_reference"] = {"version": ref["version"], "triton": ref["triton"], "gpu": ref["gpu"], "rc": rc}
    yield ref, tensors
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "vs_reference.json"), "w") as f:
        json.dump(REPORT, f, indent=1, default=str)


def _oracle_rows(layer, case, M, i, rows):
    p = oracle_from_layer(layer)
    x = refdual.x_for(case, M, i).cpu()
    return O.forward(p, x[rows]).double()


def _exact_rows(layer, x, rows):
    """x[rows] @ W with W = the stored parameters dequantised in fp64 WITHOUT rounding (A16W4 grouped INT layers)."""
    from gemlite_b200 import bitpack as B
    K, N, gs = layer.in_features, layer.out_features, layer.group_size
    codes = B.unpack_over_cols(layer.W_q.data, layer.W_nbits, K, transpose=True).double()        # (N, K)
    s = layer.scales.data.double().t().repeat_interleave(gs, dim=1)                              # (N, K)
    mode = layer.W_group_mode
    if mode == 2:
        W = codes * s
    else:
        z = layer.zeros.data.double().t().repeat_interleave(gs, dim=1)
        W = codes * s + z if mode == 4 else (codes - z) * s
    out = x[rows.to(x.device)].double() @ W.t()
    if layer.bias is not None:
        out = out + layer.bias.data.double()
    return out.cpu()


def _err(y, yo):
    d = (y.double() - yo).abs()
    return d.max().item() / (yo.abs().max().item() + 1e-30), d.mean().item()


@pytest.mark.parametrize("name", [c["name"] for c in CASES])
def test_same_inputs_same_outputs_as_reference(both, name):
    ref_all, tensors = both
    case = CASE_BY_NAME[name]
    r = ref_all["cases"].get(name)
    assert r is not None and "fatal" not in r and "skipped" not in r, (name, r)
    layer = refdual.build_layer(gemlite_b200, case["build"], tensors)
    rep = REPORT.setdefault(name, {"calls": {}, "ref_errors": r.get("err", {}), "ref_seconds": r.get("seconds")})

    # ---- packed tensors and metadata: bit-equal to what the reference's pack() produced on this GPU
    assert [int(v) for v in layer.get_meta_args()] == r["meta_args"], (layer.get_meta_args(), r["meta_args"])
    if "W_q" in r:
        for nm, ours, theirs in zip(("W_q", "scales", "zeros"), layer.get_tensor_args(), (r["W_q"], r["scales"], r["zeros"])):
            ours = ours.detach().cpu()
            assert ours.dtype == theirs.dtype and ours.shape == theirs.shape and ours.stride() == theirs.stride(), nm
            if ours.numel():
                a, b = ours.contiguous(), theirs.contiguous()
                assert torch.equal(a.view(torch.uint8) if a.dtype.is_floating_point else a,
                                   b.view(torch.uint8) if b.dtype.is_floating_point else b), f"{name}: packed {nm} differs"
        rep["packed_equal"] = True

    is_int8 = case["build"].get("cls") == "A8W8_INT8_dynamic"
    out_dt = refdual.TD[case["build"].get("dtype", "float16")]
    ulp = 2.0 ** -10 if out_dt == torch.float16 else 2.0 ** -7
    for i, M in enumerate(case["Ms"]):
        x = refdual.x_for(case, M, i)
        rows = torch.arange(M) if M <= case.get("sample_rows", 10 ** 9) else \
            torch.randperm(M, generator=torch.Generator().manual_seed(M))[:case["sample_rows"]].sort().values
        yo = yx = None
        for k in case["kernels"]:
            if k != "AUTO" and M > 1 and "GEMV" in k:
                continue
            key = f"{k}:{M}"
            y = (layer(x) if k == "AUTO" else layer.forward_manual(x, matmul_type=k))
            torch.cuda.synchronize()
            kern = _cabi.last_kernel()
            y = y.cpu()
            assert torch.isfinite(y.float()).all()
            if yo is None:
                yo = _oracle_rows(layer, case, M, i, rows)
            e_max, e_mean = _err(y[rows], yo)
            tol_o = (REL_DECODE if (kern == "gemv_w4_mma" or kern.endswith("_packed")) else REL)[out_dt]
            if layer.scaled_activations and not is_int8:
                tol_o = max(tol_o, 2e-2)       # a dynamically quantised activation code on an exact tie (see test_gpu_parity)
            assert e_max <= tol_o, (name, key, kern, "vs oracle", e_max, tol_o)
            if key not in r["y"]:
                # the reference itself raised on this GPU (recorded): the oracle comparison above is all there is
                rep["calls"][key] = dict(kernel=kern, ours_vs_oracle=(e_max, e_mean), reference_failed=r["err"].get(key, "missing"))
                continue
            yr = r["y"][key]
            assert y.dtype == yr.dtype and y.shape == yr.shape
            r_max, r_mean = _err(yr[rows], yo)
            d_mean = (y.double() - yr.double()).abs().mean().item()
            equal = bool(torch.equal(y, yr))
            call = dict(kernel=kern, ours_vs_oracle=(e_max, e_mean), ref_vs_oracle=(r_max, r_mean),
                        mean_abs_ours_minus_ref=d_mean, bit_equal=equal)
            rep["calls"][key] = call
            if is_int8 and k in case.get("exact", ()) and (k != "GEMV_SPLITK" or M == 1):
                # INT8 x INT8 -> INT32 is exact; epilogue cvt.rn.f32.s32(acc) * (s_x * s_w) -> fp16 is one rounding
                assert equal, (name, key, kern, (y != yr).sum().item(), d_mean)
            if d_mean >= 1e-3:
                # the reference's own test bar (mean |diff| < 1e-3) is missed: acceptable only if the REFERENCE is the
                # outlier w.r.t. the fp32 oracle (its GEMV kernels accumulate in fp16 with atomics; its default
                # GEMV_SPLITK / GEMV_REVSPLITK configs read one scale row per BLOCK_K and break on grouped weights)
                call["reference_outlier"] = True
                ours = e_mean
                if kern.endswith("_packed"):       # factorised kernel: its own error is the one against the exact product
                    if yx is None:
                        yx = _exact_rows(layer, x, rows)
                    ours = min(e_mean, _err(y[rows], yx)[1])
                # 1e-4 absolute + half an output ulp at the mean output magnitude (fp16 outputs around 4 are rounded by ~5e-4)
                assert ours < 1e-4 + 0.5 * ulp * max(1.0, yo.abs().mean().item()) and r_mean > 0.5 * d_mean, (name, key, kern, d_mean, ours, r_mean)
                continue
            if kern == "gemv_w4_mma" or kern.endswith("_packed"):      # factorised decode kernels (4-bit; packed 8 / 2 / 1-bit)
                # The decode kernel applies scale / zero per group to fp32 partial sums instead of rounding every
                # dequantised weight to fp16 first, so it deviates from the fp16-dequant oracle by weight-rounding noise
                # while being CLOSER to the exact product of the stored parameters.  Judge both implementations against
                # that exact product (fp64, no intermediate rounding).
                if yx is None:
                    yx = _exact_rows(layer, x, rows)
                ex_max, ex_mean = _err(y[rows], yx)
                rx_max, rx_mean = _err(yr[rows], yx)
                call["vs_exact_product"] = dict(ours=(ex_max, ex_mean), ref=(rx_max, rx_mean))
                assert ex_max <= rx_max * 1.05 + 2 * ulp, (name, key, kern, "max vs exact", ex_max, rx_max)
                assert ex_mean <= rx_mean * 1.05 + 1e-5, (name, key, kern, "mean vs exact", ex_mean, rx_mean)
            else:
                assert e_max <= r_max * 1.05 + 2 * ulp, (name, key, kern, "max", e_max, r_max)
                assert e_mean <= r_mean * 1.05 + 5e-5, (name, key, kern, "mean", e_mean, r_mean)
    if case.get("actq"):
        for i, M in enumerate(case["Ms"]):
            x = refdual.x_for(case, M, i)
            q, s = gemlite_b200.quant_utils.scale_activations_per_token(x, w_dtype=case["actq"])
            qr, sr = r["actq"][M]
            same_q = (q.cpu().view(torch.uint8) == qr.view(torch.uint8))
            rep["calls"][f"actq:{M}"] = dict(codes_equal=float(same_q.float().mean()), scales_equal=bool(torch.equal(s.cpu(), sr)))
            assert torch.equal(s.cpu().view(-1), sr.view(-1)), (name, M, "activation scales")
            assert same_q.all(), (name, M, "activation codes differ", int((~same_q).sum()))
