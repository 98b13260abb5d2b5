"""One launch of each secondary kernel on its named config (for ncu captures): A8W8 int8 / fp8 at M=4096 (quantiser +
CTA-pair GEMM), MXFP4 decode at M=1 and MXFP4 prefill at M=4096 on 16384^2, A16W4 split-K at M=64 on 4096^2."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemlite_b200.helper import A16W4_HQQ_INT, A16W4_MXFP, A8W8_FP8_dynamic, A8W8_INT8_dynamic
DEV = torch.device("cuda:0")
g = torch.Generator(device=DEV).manual_seed(3)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "a8"):
    K = N = 4096
    W = (torch.randn(N, K, generator=g, device=DEV) / 10).half()
    x = (torch.randn(4096, K, generator=g, device=DEV) / 10).half()
    for proc in (A8W8_INT8_dynamic, A8W8_FP8_dynamic):
        layer = proc(device=DEV, dtype=torch.float16).from_weights(W.clone())
        for _ in range(2):
            layer(x)
        torch.cuda.synchronize()
if which in ("all", "mx"):
    K = N = 16384
    W_q = torch.randint(0, 16, (N, K), generator=g, device=DEV, dtype=torch.uint8)
    sc = torch.randint(118, 126, (N, K // 32), generator=g, device=DEV, dtype=torch.uint8).view(torch.float8_e8m0fnu)
    layer = A16W4_MXFP(device=DEV, dtype=torch.bfloat16).from_weights(W_q, sc)
    del W_q
    for M in (1, 4096):
        x = (torch.randn(M, K, generator=g, device=DEV) / 10).bfloat16()
        for _ in range(2):
            layer(x)
        torch.cuda.synchronize()
if which in ("all", "splitk"):
    K = N = 4096
    W_q = torch.randint(0, 16, (N, K), generator=g, device=DEV, dtype=torch.int32).to(torch.uint8)
    s = (torch.rand(N * K // 128, 1, generator=g, device=DEV) * 0.01 + 0.001).half()
    z = (torch.rand(N * K // 128, 1, generator=g, device=DEV) * 15).half()
    layer = A16W4_HQQ_INT(device=DEV, dtype=torch.float16).from_weights(W_q, s, z)
    x = (torch.randn(64, K, generator=g, device=DEV) / 10).half()
    for _ in range(2):
        layer(x)
    torch.cuda.synchronize()
