"""Measured 8-bit tensor-core ceilings on this box (do not assume 2x the bf16 peak): cuBLASLt through
torch._scaled_mm (e4m3 x e4m3 -> bf16) and torch._int_mm (int8 x int8 -> int32) at 8192^3, best of 10 and sustained,
next to the bf16 matmul of MEASURED_PEAKS.json's recipe.  Library calls, used only as yardsticks.

    python tools/measure_peaks.py > profiles/tensor_peaks_r2.json
"""
import json
import torch

dev = "cuda:0"
N = 8192


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        fn()
    b.record(); torch.cuda.synchronize()
    return best, a.elapsed_time(b) / 40


out = {"shape": f"{N}^3", "flops": 2.0 * N ** 3}
x = torch.randn(N, N, device=dev, dtype=torch.bfloat16)
w = torch.randn(N, N, device=dev, dtype=torch.bfloat16)
b, s = timed(lambda: torch.matmul(x, w))
out["bf16_tflops"] = {"burst": out["flops"] / b / 1e9, "sustained": out["flops"] / s / 1e9}
try:
    xf = (torch.randn(N, N, device=dev) / 4).to(torch.float8_e4m3fn)
    wf = (torch.randn(N, N, device=dev) / 4).to(torch.float8_e4m3fn).t()           # column-major B
    one = torch.ones((), device=dev)
    b, s = timed(lambda: torch._scaled_mm(xf, wf, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
    out["fp8_e4m3_tflops"] = {"burst": out["flops"] / b / 1e9, "sustained": out["flops"] / s / 1e9}
except Exception as e:
    out["fp8_e4m3_tflops"] = {"error": repr(e)[:200]}
try:
    xi = torch.randint(-127, 127, (N, N), device=dev, dtype=torch.int8)
    wi = torch.randint(-127, 127, (N, N), device=dev, dtype=torch.int8).t()
    b, s = timed(lambda: torch._int_mm(xi, wi))
    out["int8_tops"] = {"burst": out["flops"] / b / 1e9, "sustained": out["flops"] / s / 1e9}
except Exception as e:
    out["int8_tops"] = {"error": repr(e)[:200]}
print(json.dumps(out, indent=1))
