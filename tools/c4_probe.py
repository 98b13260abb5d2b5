import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemlite_b200.helper import A8W8_FP8_dynamic, A8W8_INT8_dynamic
DEV = torch.device("cuda:0")
K = N = 4096
g = torch.Generator(device=DEV).manual_seed(3)
W = (torch.randn(N, K, generator=g, device=DEV) / 10).half()
for proc in (A8W8_FP8_dynamic, A8W8_INT8_dynamic):
    layer = proc(device=DEV, dtype=torch.float16).from_weights(W.clone())
    for M in (4096, 1024):
        x = (torch.randn(M, K, generator=g, device=DEV) / 10).half()
        for _ in range(3):
            layer(x)
        torch.cuda.synchronize()
