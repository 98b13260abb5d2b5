"""One prefill call (A16W4 gs=128, M=4096, K=N=8192) + parity check vs torch fp32; used under ncu and stand-alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gemlite_b200 import _cabi
from gemlite_b200.helper import A16W4_HQQ_INT

dev = "cuda:0"
M, K, N, GS = int(os.environ.get("PF_M", 4096)), int(os.environ.get("PF_K", 8192)), int(os.environ.get("PF_N", 8192)), 128
g = torch.Generator(device=dev).manual_seed(1)
W_q = torch.randint(0, 16, (N, K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
scales = (torch.rand(N * K // GS, 1, generator=g, device=dev) * 0.01 + 0.001).half()
zeros = (torch.rand(N * K // GS, 1, generator=g, device=dev) * 15).half()
layer = A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros)
x = (torch.randn(M, K, generator=g, device=dev) / 10).half()
y = layer(x)
torch.cuda.synchronize()
print("kernel:", _cabi.last_kernel())
# reference: exact reference-rounded weights (fp16 fma) then fp32 matmul in two halves to bound memory
Wd = torch.addcmul(layer.zeros.data.repeat_interleave(GS, 0), W_q.t().to(torch.float16), layer.scales.data.repeat_interleave(GS, 0))  # (K,N) fp16, single-rounding fma? (addcmul in fp16 computes in fp32 then rounds once)
yref = (x.float() @ Wd.float())
err = ((y.float() - yref).abs().max() / yref.abs().max()).item()
print(f"rel_err={err:.3e}")
ts = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    layer(x)
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); layer(x); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
print(f"M={M} K={K} N={N}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
assert err < 2e-3, err
