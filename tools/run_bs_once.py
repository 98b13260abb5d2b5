"""One call of a dynamic MX / NV layer (for ncu): python tools/run_bs_once.py A8W8_MXFP_dynamic 4096 4096 4096"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gemlite_b200 import helper as H

cls, N, K, M = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
kw = {"post_scale": False} if cls.startswith("A8") else {}
lin = torch.nn.Linear(K, N, bias=False, dtype=torch.bfloat16, device="cuda:0")
lin.weight.requires_grad = False
layer = getattr(H, cls)(device="cuda:0", dtype=torch.bfloat16, **kw).from_linear(lin)
x = (torch.randn(M, K, device="cuda:0") / 10).to(torch.bfloat16)
for _ in range(3):
    y = layer(x)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
