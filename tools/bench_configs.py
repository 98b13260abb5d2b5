#!/usr/bin/env python
"""Measured table for the BASELINE.json configs that are not the bench.py headline:

  C3  A16W4 gs=128 prefill  M=4096           K=N=8192 (and 4096^2)
  C4  A8W8 dynamic (fp8 x fp8 and int8 x int8), M sweep 1..4096, K=N=4096   -- includes the activation quantiser
  C5  A16W4 MXFP4 gs=32 weight-only, K=N=16384, M in {1, 16, 4096}           (single GPU here; sharded: bench.py --gpus)
  Cx  A16W4 gs=128 small-batch sweep M=1..256, K=N=4096                       (decode <-> prefill transition)

Every line: cold-L2 median of CUDA-event times of single calls (256 MB flush between calls; includes the launch
latency of an isolated call), and for M <= 64 additionally `steady=`: the sustained time per call over a rotating set of
weight copies larger than L2 inside one CUDA graph (SURVEY 8d asks for both), the algorithmic bytes / flops of
SURVEY.md 8d, the time floor max(bytes / measured HBM peak, flops / tensor peak) and the fraction floor / measured.
Tensor peaks: bf16 = MEASURED_PEAKS.json; fp8 / int8 = 2 x that (nominal ratio, no measured number on this pool).

    python tools/bench_configs.py [--quick] [--small] [--only=C3,C4,...]  > profiles/configs_rN.txt     (--small: C4 at M <= 32 only)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gemlite_b200 import _cabi  # noqa: E402
from gemlite_b200.helper import (A16W4_HQQ_INT, A16W4_MXFP, A8W8_FP8_dynamic, A8W8_INT8_dynamic)  # noqa: E402

DEV = torch.device("cuda:0")
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}
HBM = PEAKS.get("hbm_gbs", 6650.0) * 1e9
TC16 = PEAKS.get("bf16_tflops", 1590.0) * 1e12
FLUSH = None


def timed(fn, iters):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def timed_steady(layer, x, weight_bytes, replays=20):
    """Sustained time per call: enough deep copies of the layer that one pass streams > 200 MB of distinct weights
    (> the 126 MB L2), every copy called once with the same x inside ONE CUDA graph (launch overhead excluded the same
    way for every contender; consecutive calls are independent, so weight prefetch under programmatic dependent launch
    overlaps as it does in a serving loop).  SURVEY 8d asks for this next to the flushed single-call number."""
    import copy
    n = max(2, min(48, int(200e6 // max(weight_bytes, 1)) + 1))
    copies = [layer] + [copy.deepcopy(layer) for _ in range(n - 1)]
    stream = torch.cuda.Stream(device=DEV)
    keep = []
    with torch.cuda.stream(stream):
        for c in copies:
            c(x)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for c in copies:
                keep.append(c(x))
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    del graph, keep, copies
    return a.elapsed_time(b) * 1e-3 / (replays * n)


def steady_if_small(layer, x, M, weight_bytes):
    if M > 64 or "--no-steady" in sys.argv:
        return None
    try:
        return timed_steady(layer, x, weight_bytes)
    except Exception as e:   # a diagnostic column must not lose the line
        print(f"# steady-state timing failed: {e!r}"[:200], flush=True)
        return None


def line(tag, M, K, N, t, nbytes, flops, tc_peak, kern, steady=None):
    floor_b, floor_f = nbytes / HBM, flops / tc_peak
    floor = max(floor_b, floor_f)
    bound = "hbm" if floor_b >= floor_f else "tensor"
    extra = "" if steady is None else f"  steady={steady * 1e6:7.2f} us frac={floor / steady:5.3f}"
    print(f"{tag:26s} M={M:5d} K={K:5d} N={N:5d}  {t * 1e6:9.1f} us  {nbytes / t / 1e9:8.1f} GB/s  {flops / t / 1e12:8.1f} TFLOP/s  "
          f"bound={bound:6s} floor={floor * 1e6:8.1f} us  frac={floor / t:6.3f}  kernel={kern}{extra}", flush=True)


def c3(quick):
    for (M, K, N) in ([(4096, 8192, 8192)] if quick else [(4096, 8192, 8192), (4096, 4096, 4096)]):
        g = torch.Generator(device=DEV).manual_seed(1)
        W_q = torch.randint(0, 16, (N, K), generator=g, device=DEV, dtype=torch.int32).to(torch.uint8)
        ng = N * K // 128
        s = (torch.rand(ng, 1, generator=g, device=DEV) * 0.01 + 0.001).half()
        z = (torch.rand(ng, 1, generator=g, device=DEV) * 15).half()
        layer = A16W4_HQQ_INT(device=DEV, dtype=torch.float16).from_weights(W_q, s, z)
        del W_q
        x = (torch.randn(M, K, generator=g, device=DEV) / 10).half()
        t = timed(lambda: layer(x), 5 if quick else 10)
        nbytes = K * N // 2 + 2 * (K // 128) * N * 2 + 2 * M * K + 2 * M * N
        line("C3 A16W4 gs128 prefill", M, K, N, t, nbytes, 2.0 * M * K * N, TC16, _cabi.last_kernel())
        del layer, x


def small_batch(quick):
    K = N = 4096
    g = torch.Generator(device=DEV).manual_seed(2)
    W_q = torch.randint(0, 16, (N, K), generator=g, device=DEV, dtype=torch.int32).to(torch.uint8)
    ng = N * K // 128
    s = (torch.rand(ng, 1, generator=g, device=DEV) * 0.01 + 0.001).half()
    z = (torch.rand(ng, 1, generator=g, device=DEV) * 15).half()
    layer = A16W4_HQQ_INT(device=DEV, dtype=torch.float16).from_weights(W_q, s, z)
    del W_q
    for M in ([1, 16, 64] if quick else [1, 2, 4, 8, 16, 32, 64, 128, 256]):
        x = (torch.randn(M, K, generator=g, device=DEV) / 10).half()
        t = timed(lambda: layer(x), 5 if quick else 20)
        nbytes = K * N // 2 + 2 * (K // 128) * N * 2 + 2 * M * K + 2 * M * N
        kern = _cabi.last_kernel()
        line("Cx A16W4 gs128 small-M", M, K, N, t, nbytes, 2.0 * M * K * N, TC16, kern, steady_if_small(layer, x, M, K * N // 2))


def c4(quick):
    K = N = 4096
    g = torch.Generator(device=DEV).manual_seed(3)
    W = (torch.randn(N, K, generator=g, device=DEV) / 10).half()
    for name, proc in (("C4 A8W8 fp8 dynamic", A8W8_FP8_dynamic), ("C4 A8W8 int8 dynamic", A8W8_INT8_dynamic)):
        layer = proc(device=DEV, dtype=torch.float16).from_weights(W.clone())
        for M in ([32, 64, 128] if "--mid" in sys.argv else [1, 2, 4, 8, 16, 32] if "--small" in sys.argv else ([1, 64, 4096] if quick else [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])):
            x = (torch.randn(M, K, generator=g, device=DEV) / 10).half()
            t = timed(lambda: layer(x), 5 if quick else 20)
            # W K*N + 4N, x_q M*K + 4M, y 2*M*N, plus the quantiser pass (2*M*K read, M*K write)
            nbytes = K * N + 4 * N + M * K + 4 * M + 2 * M * N + 3 * M * K
            kern = _cabi.last_kernel()
            line(name, M, K, N, t, nbytes, 2.0 * M * K * N, 2 * TC16, kern, steady_if_small(layer, x, M, K * N))
        del layer


def c5(quick):
    K = N = 16384
    g = torch.Generator(device=DEV).manual_seed(4)
    # from_weights takes one e2m1 code per byte (N, K) and packs two per byte itself (helper.py: A16Wn_MXFP)
    W_q = torch.randint(0, 16, (N, K), generator=g, device=DEV, dtype=torch.uint8)
    scales = torch.randint(118, 126, (N, K // 32), generator=g, device=DEV, dtype=torch.uint8).view(torch.float8_e8m0fnu)
    layer = A16W4_MXFP(device=DEV, dtype=torch.bfloat16).from_weights(W_q, scales)
    del W_q
    for M in ([1, 16] if quick else [1, 16, 4096]):
        x = (torch.randn(M, K, generator=g, device=DEV) / 10).bfloat16()
        t = timed(lambda: layer(x), 3 if M > 64 else (5 if quick else 10))
        nbytes = K * N // 2 + K * N // 32 + 2 * M * K + 2 * M * N
        kern = _cabi.last_kernel()
        line("C5 A16W4 MXFP4 gs32", M, K, N, t, nbytes, 2.0 * M * K * N, TC16, kern, steady_if_small(layer, x, M, K * N // 2))


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    only = None
    for a in sys.argv[1:]:
        if a.startswith("--only="):
            only = set(a[7:].split(","))
    print(f"# peaks: HBM {HBM / 1e9:.1f} GB/s, bf16 tensor {TC16 / 1e12:.1f} TFLOP/s (MEASURED_PEAKS.json), fp8/int8 tensor = 2x bf16 (nominal ratio)")
    for tag, fn in (("C3", c3), ("Cx", small_batch), ("C4", c4), ("C5", c5)):
        if only is None or tag in only:
            try:
                fn(quick)
            except Exception as e:  # keep the table going
                print(f"{tag}: FAILED {e!r}"[:300], flush=True)
