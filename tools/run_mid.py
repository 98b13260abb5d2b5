"""Run one A16W4 (or A8W8) 4096^2 layer at batch size M a few times (for ncu captures of the batched-decode kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gemlite_b200, gemlite_b200.helper
from gemlite_b200 import _cabi
import refdual
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "w4"
if kind == "w4":
    b = dict(kind="pack_int", nbits=4, gs=128, N=4096, K=4096, dtype="float16", seed=1)
else:
    b = dict(kind="helper_w", cls="A8W8_INT8_dynamic", N=4096, K=4096, dtype="float16", seed=1)
layer = refdual.build_layer(gemlite_b200, b)
x = refdual.gen_x(M, 4096, torch.float16, 3).to("cuda:0")
for _ in range(5):
    y = layer(x)
torch.cuda.synchronize()
print(_cabi.last_kernel(), float(y.float().abs().max()))
