#!/usr/bin/env python
"""Prefill sweep: A16W4 gs=128 through the public API, checked against a torch fp32 reference and timed
cold-L2 (256 MB flush) with CUDA events.  Used under ncu for the prefill capture (a few launches only).

    python tools/run_prefill.py [--quick] [--shapes M,K,N;M,K,N...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemlite_b200 import _cabi  # noqa: E402
from gemlite_b200.helper import A16W4_HQQ_INT  # noqa: E402

GROUP = 128


def run(M, K, N, dtype=torch.float16, iters=10, check=True):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    W_q = torch.randint(0, 16, (N, K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    ng = N * K // GROUP
    scales = (torch.rand(ng, 1, generator=g, device=dev) * 0.01 + 0.001).to(dtype)
    zeros = (torch.rand(ng, 1, generator=g, device=dev) * 15).to(dtype)
    layer = A16W4_HQQ_INT(device=dev, dtype=dtype).from_weights(W_q, scales, zeros)
    x = (torch.randn(M, K, generator=g, device=dev) / 10).to(dtype)
    y = layer(x)
    kern = _cabi.last_kernel()
    rel = float("nan")
    if check:
        Wd = ((W_q.view(ng, GROUP).float() - zeros.float()) * scales.float()).view(N, K)
        rows = min(M, 256)
        ref = x[:rows].float() @ Wd.t()
        rel = ((y[:rows].float() - ref).abs().max() / ref.abs().max()).item()
        del Wd, ref
    del W_q
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); layer(x); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(f"kernel={kern} rel_err={rel:.3e} M={M} K={K} N={N} {str(dtype)[6:]}: {ms:.3f} ms  {2.0 * M * K * N / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    shapes = [(4096, 8192, 8192)] if quick else [(4096, 8192, 8192), (4096, 4096, 4096), (1024, 8192, 8192), (128, 8192, 8192), (32, 4096, 4096)]
    for a in sys.argv[1:]:
        if a.startswith("--shapes="):
            shapes = [tuple(int(v) for v in s.split(",")) for s in a[9:].split(";")]
    for (M, K, N) in shapes:
        run(M, K, N, iters=3 if quick else 10)
