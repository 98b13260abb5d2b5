"""One case list, two implementations: the same construction + forward code is run against `gemlite_b200` (in the
test process) and against the UNMODIFIED reference package `gemlite` (installed under baseline/_ref, executed in a
child process by tools/ref_runner.py -- both packages register the torch op `gemlite::forward_functional`, so they
cannot share an interpreter).  Both sides import THIS file, so "identical inputs" holds by construction: tensors come
from seeded CPU generators (or are shipped explicitly in the job file) and every call below is part of the public API
the two packages share (GemLiteLinear / pack / forward_manual / helper processors / quant_utils).

SURVEY 8(c): the live reference on the GPU box is the strongest oracle there is; tests/test_gpu_vs_reference.py is
skipped only when baseline/_ref is absent (it is git-ignored but travels with gpurun)."""
import time

import torch

DEV = "cuda:0"
TD = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}
KERNELS = ["GEMV_REVSPLITK", "GEMV", "GEMV_SPLITK", "GEMM_SPLITK", "GEMM"]


def gen_int(N, K, nbits, gs, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    W_q = torch.randint(0, 2 ** nbits, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    ng = N * K // gs
    scales = (torch.rand(ng, 1, generator=g) * 0.01 + 0.001).to(dtype)
    zeros = (torch.rand(ng, 1, generator=g) * (2 ** nbits - 1)).to(dtype)
    return W_q, scales, zeros


def gen_w(N, K, dtype, seed):
    return (torch.randn(N, K, generator=torch.Generator().manual_seed(seed)) / 10.0).to(dtype)


def gen_x(M, K, dtype, seed):
    return (torch.randn(M, K, generator=torch.Generator().manual_seed(seed)) / 10.0).to(dtype)


# ----------------------------------------------------------------------------------------------------------------
# layer construction through the shared public API.  `pkg` is the imported package (gemlite or gemlite_b200).
# ----------------------------------------------------------------------------------------------------------------
def build_layer(pkg, b, tensors=None):
    kind = b["kind"]
    dt = TD[b.get("dtype", "float16")]
    H = pkg.helper
    if kind == "pack_int":                       # GemLiteLinear(...).pack(W_q, scales, zeros): core.py:231-519
        W_q, s, z = gen_int(b["N"], b["K"], b["nbits"], b["gs"], dt, b["seed"])
        gd = pkg.dtypes.TORCH_TO_DTYPE[dt]
        layer = pkg.GemLiteLinear(b["nbits"], b["gs"], b["K"], b["N"], gd, gd)
        bias = gen_w(1, b["N"], dt, b["seed"] + 5).view(-1).to(DEV) if b.get("bias") else None
        zz = z.to(DEV) if not b.get("symmetric") else None
        if b.get("scalar_zero") is not None:
            zz = b["scalar_zero"]
        layer.pack(W_q.to(DEV), s.to(DEV), zz, bias, fma_mode=b.get("fma_mode", True),
                   packing_bitwidth=b.get("packing_bitwidth"))
        return layer
    if kind == "fp16":                           # W_nbits=16: test_gemlitelineartriton.py:87-113
        W = gen_w(b["N"], b["K"], dt, b["seed"])
        gd = pkg.dtypes.TORCH_TO_DTYPE[dt]
        layer = pkg.GemLiteLinear(W_nbits=16, group_size=None, in_features=b["K"], out_features=b["N"],
                                  input_dtype=gd, output_dtype=gd, scaled_activations=False)
        layer.pack(W.to(DEV), None, None, None)
        return layer
    if kind == "helper_w":                       # processors that quantise a float weight themselves
        W = gen_w(b["N"], b["K"], dt, b["seed"]).to(DEV)
        proc = getattr(H, b["cls"])(device=DEV, dtype=dt, **b.get("kwargs", {}))
        return proc.from_weights(W)
    if kind == "hqq_int":                        # A16Wn_HQQ_INT.from_weights(W_q, scales, zeros): helper.py:344-400
        W_q, s, z = gen_int(b["N"], b["K"], b["nbits"], b["gs"], dt, b["seed"])
        proc = getattr(H, b["cls"])(device=DEV, dtype=dt)
        return proc.from_weights(W_q.to(DEV), s.to(DEV), z.to(DEV))
    if kind == "mx_weights":                     # pre-quantised MX codes + e8m0 scales shipped in the job
        t = tensors[b["tensors"]]
        proc = getattr(H, b["cls"])(device=DEV, dtype=dt)
        return proc.from_weights(t["W_q"].to(DEV), t["scales"].to(DEV))
    if kind == "mx_random":                      # timing only: random e2m1 codes / e8m0 scales made on the device
        g = torch.Generator(device=DEV).manual_seed(b["seed"])
        W_q = torch.randint(0, 16, (b["N"], b["K"]), generator=g, device=DEV, dtype=torch.uint8)
        sc = torch.randint(118, 126, (b["N"], b["K"] // 32), generator=g, device=DEV, dtype=torch.uint8)
        proc = getattr(H, b["cls"])(device=DEV, dtype=dt)
        return proc.from_weights(W_q, sc.view(torch.float8_e8m0fnu))
    if kind == "torch_matmul":                   # the reference's yardstick (examples/benchmark_triton.py:218)
        W = gen_w(b["N"], b["K"], dt, b["seed"]).to(DEV)
        return lambda x, Wt=W.t(): torch.matmul(x, Wt)
    if kind == "mx_dynamic":                     # A8W8/A8W4/A4W4 MXFP/NVFP dynamic: helper.py:660-945
        W = gen_w(b["N"], b["K"], dt, b["seed"]).to(DEV)
        lin = torch.nn.Linear(b["K"], b["N"], bias=False, dtype=dt, device=DEV)
        lin.weight.data = W
        lin.weight.requires_grad = False
        proc = getattr(H, b["cls"])(device=DEV, dtype=dt, **b.get("kwargs", {}))
        return proc.from_linear(lin)
    if kind == "bitnet":
        g = torch.Generator().manual_seed(b["seed"])
        W = torch.randint(-1, 2, (b["N"], b["K"]), generator=g).to(dt)
        sc = torch.rand(1, generator=g).to(dt) * 0 + 0.5
        proc = getattr(H, b["cls"])(device=DEV, dtype=dt)
        return proc.from_weights(W.to(DEV), sc.to(DEV))
    raise ValueError(kind)


def x_for(case, M, i):
    b = case["build"]
    dt = TD[b.get("dtype", "float16")]
    return gen_x(M, b["K"], dt, 1000 + 17 * i + M).to(DEV)


def run_case(pkg, case, tensors=None):
    """-> dict: packed tensors + meta + outputs per (kernel, M). Exceptions are recorded per call, not raised."""
    layer = build_layer(pkg, case["build"], tensors)
    out = {"meta_args": [int(v) for v in layer.get_meta_args()], "y": {}, "err": {}}
    if case.get("keep_packed", True):
        ta = layer.get_tensor_args()
        out["W_q"], out["scales"], out["zeros"] = [t.detach().cpu() if t is not None else None for t in ta]
    for i, M in enumerate(case["Ms"]):
        x = x_for(case, M, i)
        for k in case["kernels"]:
            if k != "AUTO" and M > 1 and "GEMV" in k:
                continue
            try:
                y = layer(x) if k == "AUTO" else layer.forward_manual(x, matmul_type=k)
                torch.cuda.synchronize()
                out["y"][f"{k}:{M}"] = y.detach().cpu()
            except Exception as e:  # noqa: BLE001 -- recorded and asserted on by the test
                out["err"][f"{k}:{M}"] = repr(e)[:300]
        del x
    if case.get("actq"):                         # per-token activation quantiser on its own (quant_utils.py:268-347)
        out["actq"] = {}
        for i, M in enumerate(case["Ms"]):
            x = x_for(case, M, i)
            q, s = pkg.quant_utils.scale_activations_per_token(x, w_dtype=case["actq"])
            out["actq"][M] = (q.detach().cpu(), s.detach().cpu())
    del layer
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------
# timing protocols, shared by both contenders
# ----------------------------------------------------------------------------------------------------------------
def time_flush(fn, rep=200, cache=None):
    """The reference's own protocol (examples/benchmark_triton.py:44-60): 256 MB scratch zeroed before every timed
    call, one event pair per call, host launch overhead included.  Returns (min_ms, median_ms)."""
    import random
    if cache is None:
        cache = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int, device=DEV)
    ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cache.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        cache += int(random.random() * 1000)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def time_graph_rotating(layers, x, chained, iters=20, warm=3):
    """Launch overhead excluded the same way for both: every layer of `layers` (distinct weights, together larger than
    L2) once per step, captured in ONE CUDA graph; chained=True feeds y of layer i to layer i+1 (needs N == K).
    Returns ms per layer call (device time of the replay / len(layers))."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            h = x
            for L in layers:
                y = L(h)
                h = y if chained else x
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            h = x
            for L in layers:
                y = L(h)
                h = y if chained else x
        for _ in range(warm):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / len(layers)


def time_case(pkg, t):
    """t: {build, M, n_layers, kernel, chained, rep} -> dict of timings (ms per call)."""
    b = dict(t["build"])
    n_layers = t.get("n_layers", 1)
    layers = []
    for i in range(n_layers):
        bb = dict(b)
        bb["seed"] = b["seed"] + i
        layers.append(build_layer(pkg, bb))
    k = t.get("kernel", "AUTO")
    calls = [(L if k == "AUTO" else (lambda h, L=L: L.forward_manual(h, matmul_type=k))) for L in layers]
    dt = TD[b.get("dtype", "float16")]
    x = gen_x(t["M"], b["K"], dt, 77).to(DEV)
    t0 = time.time()
    for c in calls[:1]:
        for _ in range(3):
            c(x)                                  # the reference autotunes here
    torch.cuda.synchronize()
    res = {"first_call_s": time.time() - t0}
    for c in calls:
        c(x)
    torch.cuda.synchronize()
    mn, med = time_flush(lambda: calls[0](x), rep=t.get("rep", 200))
    res["flush_min_ms"], res["flush_median_ms"] = mn, med
    if n_layers > 1:
        try:
            res["graph_ms_per_layer"] = time_graph_rotating(calls, x, chained=False)
            if t.get("chained") and b["K"] == b["N"] and t["M"] >= 1:
                res["graph_chained_ms_per_layer"] = time_graph_rotating(calls, x, chained=True)
        except Exception as e:  # noqa: BLE001
            res["graph_error"] = repr(e)[:200]
    del layers, calls
    torch.cuda.empty_cache()
    return res
