"""GEMM-only time of two dynamic MX layers under the GLB_BS_DEBUG switches (diagnostics; results are wrong with them)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, torch
sys.path.insert(0, %r)
from gemlite_b200 import helper as H, _cabi
from gemlite_b200.quant_utils import scale_activations_mxfp8, scale_activations_nvfp4
def t(fn, rep=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / rep * 1e3
for cls, kw, q, N in (("A8W8_MXFP_dynamic", {"post_scale": False}, scale_activations_mxfp8, 4096), ("A4W4_NVFP_dynamic", {}, scale_activations_nvfp4, 8192)):
    lin = torch.nn.Linear(N, N, bias=False, dtype=torch.bfloat16, device="cuda:0"); lin.weight.requires_grad = False
    layer = getattr(H, cls)(device="cuda:0", dtype=torch.bfloat16, **kw).from_linear(lin)
    x = (torch.randn(4096, N, device="cuda:0") / 10).to(torch.bfloat16)
    call, quant = t(lambda: layer(x)), t(lambda: q(x))
    print(os.environ.get("GLB_BS_DEBUG", "0"), cls, N, "gemm_us", round(call - quant, 1), "TF", round(2 * 4096 * N * N / ((call - quant) * 1e-6) / 1e12))
''' % ROOT
for dbg in ("0", "1", "2", "3", "4", "5"):
    env = dict(os.environ, GLB_BS_DEBUG=dbg)
    subprocess.run([sys.executable, "-c", CODE], env=env, timeout=200)
