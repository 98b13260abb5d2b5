"""Diagnostic: pipelined timeline of the decode kernel (GLB_GEMV_DEBUG=4) inside a CUDA-graph replay."""
import ctypes, os, sys
os.environ["GLB_GEMV_DEBUG"] = "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gemlite_b200 import _cabi
from gemlite_b200.helper import A16W4_HQQ_INT
dev = "cuda:0"; K = N = 4096; L = 24
g = torch.Generator(device=dev).manual_seed(0)
layers = []
for _ in range(L):
    W_q = torch.randint(0, 16, (N, K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    s = (torch.rand(N * K // 128, 1, generator=g, device=dev) * 0.002 + 0.0024).half()
    z = (torch.rand(N * K // 128, 1, generator=g, device=dev) * 3 + 6).half()
    layers.append(A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, s, z))
x0 = (torch.randn(1, K, device=dev) / 10).half()
def step():
    x = x0
    for l in layers: x = l(x)
    return x
for _ in range(2): step()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
with torch.cuda.stream(st):
    step(); torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=st): step()
for _ in range(3): gr.replay()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8192)()
_cabi.load().glb_debug_timeline(buf, 8192)       # clear
gr.replay(); torch.cuda.synchronize()
n = _cabi.load().glb_debug_timeline(buf, 8192)
rec = np.array(buf[:n], dtype=np.uint64).reshape(-1, 9).astype(np.int64)
rec = rec[np.argsort(rec[:, 1])]
t0 = rec[0, 1]
names = ["start", "tma_issued", "wait_done", "staged", "meta", "w0_ready", "loop_done", "reduced"]
if os.environ.get("GLB_GEMV_VARIANT", "7") == "7":
    names = ["start", "issued", "meta_ok", "handoff", "x_staged", "loop_done", "published", "(flagged)"]
    rec[:, 8] = rec[:, 8] + rec[:, 1] - (rec[:, 8] * 0)      # last column is a flag, shown as start + flag
print("cta   " + " ".join(f"{n:>11s}" for n in names) + "   (ns since first start; deltas in brackets)")
for r in rec:
    t = r[1:] - t0
    print(f"{int(r[0]):4d}  " + " ".join(f"{int(v):6d}[{int(v - (t[i-1] if i else v)):4d}]"[:11].rjust(11) for i, v in enumerate(t)))
