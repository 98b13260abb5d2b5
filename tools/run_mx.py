"""Where the time of a dynamic MX / NV layer call goes: activation quantiser and block-scaled GEMM timed separately
(CUDA graph of `n` distinct layers / inputs, events on the capture stream)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemlite_b200
from gemlite_b200 import _cabi
from gemlite_b200 import helper as H
from gemlite_b200.quant_utils import scale_activations_mxfp8, scale_activations_mxfp4, scale_activations_nvfp4

DEV = "cuda:0"
CASES = [("A8W8_MXFP_dynamic", {"post_scale": False}, scale_activations_mxfp8, 4096, 4096),
         ("A8W8_MXFP_dynamic", {"post_scale": False}, scale_activations_mxfp8, 8192, 8192),
         ("A4W4_MXFP_dynamic", {}, scale_activations_mxfp4, 8192, 8192),
         ("A4W4_MXFP_dynamic", {}, scale_activations_mxfp4, 16384, 16384),
         ("A4W4_NVFP_dynamic", {}, scale_activations_nvfp4, 8192, 8192)]


def graph_time(fn, rep=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / rep


out = []
for cls, kw, quant, N, K in CASES:
    n = 4 if N <= 8192 else 1
    layers = []
    for i in range(n):
        lin = torch.nn.Linear(K, N, bias=False, dtype=torch.bfloat16, device=DEV)
        lin.weight.requires_grad = False
        layers.append(getattr(H, cls)(device=DEV, dtype=torch.bfloat16, **kw).from_linear(lin))
        del lin
    torch.cuda.empty_cache()
    for M in (64, 4096):
        xs = [(torch.randn(M, K, device=DEV) / 10).to(torch.bfloat16) for _ in range(n)]
        t_all = graph_time(lambda: [l(x) for l, x in zip(layers, xs)]) / n
        kern = _cabi.last_kernel()
        t_q = graph_time(lambda: [quant(x) for x in xs]) / n
        r = dict(layer=cls, N=N, K=K, M=M, kernel=kern, call_us=t_all * 1e3, quantiser_us=t_q * 1e3, gemm_us=(t_all - t_q) * 1e3,
                 gemm_tflops=2.0 * M * N * K / ((t_all - t_q) * 1e-3) / 1e12, call_tflops=2.0 * M * N * K / (t_all * 1e-3) / 1e12)
        out.append(r)
        print(json.dumps(r), flush=True)
    del layers
    torch.cuda.empty_cache()
