"""Throughput of gemlite_b200 next to the reference's Triton kernels on the same box (VERDICT r1 row g).

    python tools/bench_vs_reference.py [--only c2,c3] [--budget 600] [--out gpurun_out/vs_reference_timing.json]

Both contenders run tests/refdual.py::time_case on the same shapes / seeds:
  flush   the reference's own protocol (examples/benchmark_triton.py:44-60): 256 MB L2 flush before each eager call,
          min and median of `rep` calls -- includes each side's host launch path;
  graph   `n_layers` distinct layers (> L2 together) in one CUDA graph, same x into each (launch overhead excluded the
          same way for both); chained = y of layer i is x of layer i+1 (the decode bench's dependency chain).
The reference autotunes in its "fast" mode (its default), "max" where noted (no b200.json ships with it).
The fp16 torch.matmul yardstick the reference quotes its speed-ups against (benchmark_triton.py:218) is the third row."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

I4 = dict(kind="pack_int", nbits=4, gs=128, dtype="float16", seed=100)
A8I = dict(kind="helper_w", cls="A8W8_INT8_dynamic", dtype="float16", seed=200, N=4096, K=4096)
A8F = dict(kind="helper_w", cls="A8W8_FP8_dynamic", dtype="float16", seed=300, N=4096, K=4096)
W8 = dict(kind="helper_w", cls="A16W8_INT8", dtype="float16", seed=500, N=4096, K=4096)
W8F = dict(kind="helper_w", cls="A16W8_FP8", dtype="bfloat16", seed=510, N=4096, K=4096)
W2 = dict(kind="hqq_int", cls="A16W2_HQQ_INT", nbits=2, gs=64, dtype="float16", seed=520, N=4096, K=4096)
MX8D = dict(kind="mx_dynamic", cls="A8W8_MXFP_dynamic", dtype="bfloat16", seed=600, N=4096, K=4096, kwargs={"post_scale": False})
MX4D = dict(kind="mx_dynamic", cls="A4W4_MXFP_dynamic", dtype="bfloat16", seed=610, N=8192, K=8192)
NV4D = dict(kind="mx_dynamic", cls="A4W4_NVFP_dynamic", dtype="bfloat16", seed=620, N=8192, K=8192)
MX = dict(kind="mx_random", cls="A16W4_MXFP", dtype="bfloat16", seed=400, N=16384, K=16384)
TIMING = [
    dict(name="c2_decode_M1", build=dict(I4, N=4096, K=4096), M=1, n_layers=64, chained=True, rep=300,
         autotune={"GEMV_REVSPLITK": "max"}, bytes=8929280),
    dict(name="c2_M8", build=dict(I4, N=4096, K=4096), M=8, n_layers=32, rep=100),
    dict(name="c2_M32", build=dict(I4, N=4096, K=4096), M=32, n_layers=32, rep=100),
    dict(name="c2_M64", build=dict(I4, N=4096, K=4096), M=64, n_layers=32, rep=100),
    dict(name="c2_M128", build=dict(I4, N=4096, K=4096), M=128, n_layers=32, rep=100),
    dict(name="c3_prefill_4096", build=dict(I4, N=4096, K=4096), M=4096, n_layers=16, rep=50, flops=2 * 4096 ** 3),
    dict(name="c3_prefill_8192", build=dict(I4, N=8192, K=8192), M=4096, n_layers=4, rep=50, flops=2 * 4096 * 8192 ** 2),
    dict(name="c4_int8_M1", build=A8I, M=1, n_layers=16, rep=100),
    dict(name="c4_int8_M64", build=A8I, M=64, n_layers=16, rep=100),
    dict(name="c4_int8_M4096", build=A8I, M=4096, n_layers=8, rep=50, flops=2 * 4096 ** 3),
    dict(name="c4_fp8_M4096", build=A8F, M=4096, n_layers=8, rep=50, flops=2 * 4096 ** 3),
    dict(name="a16w8_int8_M1", build=W8, M=1, n_layers=16, rep=100, bytes=4096 * 4096),
    dict(name="a16w8_int8_M64", build=W8, M=64, n_layers=16, rep=100),
    dict(name="a16w8_int8_M4096", build=W8, M=4096, n_layers=8, rep=50, flops=2 * 4096 ** 3),
    dict(name="a16w8_fp8_M4096", build=W8F, M=4096, n_layers=8, rep=50, flops=2 * 4096 ** 3),
    dict(name="a16w2_hqq_M1", build=W2, M=1, n_layers=16, rep=100),
    dict(name="a16w2_hqq_M64", build=W2, M=64, n_layers=16, rep=100),
    dict(name="a16w2_hqq_M4096", build=W2, M=4096, n_layers=8, rep=50, flops=2 * 4096 ** 3),
    dict(name="mxdyn_a8w8_M64", build=MX8D, M=64, n_layers=8, rep=50),
    dict(name="mxdyn_a8w8_M4096", build=MX8D, M=4096, n_layers=8, rep=30, flops=2 * 4096 ** 3),
    dict(name="mxdyn_a4w4_M64", build=MX4D, M=64, n_layers=4, rep=50),
    dict(name="mxdyn_a4w4_M4096", build=MX4D, M=4096, n_layers=4, rep=20, flops=2 * 4096 * 8192 ** 2),
    dict(name="mxdyn_nvfp4_M4096", build=NV4D, M=4096, n_layers=4, rep=20, flops=2 * 4096 * 8192 ** 2),
    dict(name="c5_mxfp4_M1", build=MX, M=1, n_layers=1, rep=50, bytes=142606336),
    dict(name="c5_mxfp4_M4096", build=MX, M=4096, n_layers=1, rep=20, flops=2 * 4096 * 16384 ** 2),
]
YARD = [
    dict(name="c2_decode_M1", build=dict(kind="torch_matmul", dtype="float16", seed=1, N=4096, K=4096), M=1, n_layers=16, rep=300),
    dict(name="c3_prefill_4096", build=dict(kind="torch_matmul", dtype="float16", seed=1, N=4096, K=4096), M=4096, n_layers=8, rep=50),
    dict(name="c3_prefill_8192", build=dict(kind="torch_matmul", dtype="float16", seed=1, N=8192, K=8192), M=4096, n_layers=2, rep=50),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--budget", type=float, default=900.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "vs_reference_timing.json"))
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    sel = [t for t in TIMING if not a.only or any(k in t["name"] for k in a.only.split(","))]
    import torch
    import gemlite_b200
    import gemlite_b200.helper  # noqa: F401
    import refdual
    res = {"ours": {}, "reference": {}, "torch_fp16_matmul": {}}
    for t in sel:
        res["ours"][t["name"]] = refdual.time_case(gemlite_b200, t)
        print("[ours]", t["name"], res["ours"][t["name"]], flush=True)
    for t in YARD:
        if any(t["name"] == s["name"] for s in sel):
            res["torch_fp16_matmul"][t["name"]] = refdual.time_case(gemlite_b200, t)
    torch.cuda.empty_cache()
    if not a.skip_reference and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gemlite")):
        tmp = tempfile.mkdtemp(prefix="vsref_t_")
        job, out = os.path.join(tmp, "job.pt"), os.path.join(tmp, "out.pt")
        torch.save({"cases": [], "timing": sel, "autotune": "fast", "budget_s": a.budget}, job)
        env = dict(os.environ)
        env.pop("PYTHONPATH", None)
        try:
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_runner.py"), job, out], env=env, cwd=tmp,
                           timeout=a.budget + 300)
        except subprocess.TimeoutExpired:
            print("reference child timed out; partial results kept")
        if os.path.exists(out):
            r = torch.load(out, weights_only=False)
            res["reference"] = r["timing"]
            res["reference_info"] = {k: r[k] for k in ("version", "triton", "gpu")}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(f"{'case':18s} {'metric':22s} {'ours':>10s} {'reference':>10s} {'ref/ours':>9s} {'fp16 mm':>10s}")
    for t in sel:
        o, r, y = res["ours"].get(t["name"], {}), res["reference"].get(t["name"], {}), res["torch_fp16_matmul"].get(t["name"], {})
        for m in ("flush_min_ms", "flush_median_ms", "graph_ms_per_layer", "graph_chained_ms_per_layer"):
            if m in o or m in r:
                ov, rv, yv = o.get(m), r.get(m), y.get(m)
                ratio = f"{rv / ov:9.2f}" if ov and rv else " " * 9
                f = lambda v: f"{v * 1e3:10.2f}" if v else " " * 10  # noqa: E731
                extra = ""
                if ov and "bytes" in t:
                    extra = f"  ours {t['bytes'] / ov / 1e6:7.0f} GB/s" + (f", ref {t['bytes'] / rv / 1e6:7.0f} GB/s" if rv else "")
                if ov and "flops" in t:
                    extra = f"  ours {t['flops'] / ov / 1e9:7.0f} TF/s" + (f", ref {t['flops'] / rv / 1e9:7.0f} TF/s" if rv else "")
                print(f"{t['name']:18s} {m:22s} {f(ov)} {f(rv)} {ratio} {f(yv)}{extra}   (us)")


if __name__ == "__main__":
    main()
