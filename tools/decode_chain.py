"""Chained decode micro-benchmark: L distinct A16W4 gs=128 KxK layers, y of layer i is x of layer i+1, one CUDA graph
(PDL on), replayed.  Prints us/layer, GB/s, the fraction of the measured HBM peak, and checks the graph result bit for bit against eager calls.  Kernel flavour knobs are read once per process
by the library (GLB_GEMV_VARIANT / GLB_GEMV_FLAGS / GLB_GEMV_TRIGGER / GLB_GEMV_DEBUG): run one process per setting.

    python tools/decode_chain.py [--layers 64] [--k 4096] [--iters 50] [--dtype fp16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gemlite_b200 import DType, GemLiteLinear, _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=64)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--gs", type=int, default=128)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    dev = "cuda:0"
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    gd = DType.FP16 if a.dtype == "fp16" else DType.BF16
    K = N = a.k
    layers = []
    g = torch.Generator(device=dev).manual_seed(0)
    for i in range(a.layers):
        W_q = torch.randint(0, 16, (N, K), generator=g, device=dev, dtype=torch.uint8)
        # scales small enough that a 64-deep chain neither overflows nor vanishes: |y| ~ |x|
        s = (torch.rand(N * K // a.gs, 1, generator=g, device=dev) * 0.002 + 0.0024).to(dt)   # std(W) * sqrt(K) ~ 1 (bench.py)
        z = (torch.rand(N * K // a.gs, 1, generator=g, device=dev) * 3 + 6).to(dt)
        layers.append(GemLiteLinear(4, a.gs, K, N, gd, gd).pack(W_q, s, z))
    x = (torch.randn(1, K, generator=g, device=dev) / 10).to(dt)
    # eager reference chain
    h = x
    for L in layers:
        h = L(h)
    torch.cuda.synchronize()
    y_eager = h.clone()
    kern = _cabi.last_kernel()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        h = x
        for L in layers:
            h = L(h)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            h = x
            for L in layers:
                h = L(h)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(h, y_eager))
        finite = bool(torch.isfinite(h.float()).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(5):
            e0.record()
            for _ in range(a.iters):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / a.iters / a.layers * 1e3)
        same_after = bool(torch.equal(h, y_eager))
    G = K // a.gs
    nbytes = K * N // 2 + 2 * G * N * 2 + 2 * K + 2 * N
    peak = 6570.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    gbs = nbytes / best / 1e3
    env = {k: os.environ[k] for k in os.environ if k.startswith("GLB_")}
    print(f"{a.tag:14s} us/layer={best:6.3f} GB/s={gbs:7.0f} frac={gbs / peak:5.3f} "
          f"equal_eager={same}/{same_after} finite={finite} kernel={kern} absmax={float(h.float().abs().max()):.3g} env={env}", flush=True)


if __name__ == "__main__":
    main()
