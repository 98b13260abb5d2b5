#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at N=1 (BASELINE.json configs[1]): A16W4 group-size-128 decode, M=1, K=N=4096, fp16.
One *step* = one decode pass through LAYERS=64 distinct layers of that shape, chained (y of layer i is x of layer
i+1), i.e. 572 MB of packed weights + metadata per step -- 4.5x the 126 MB L2, so every layer streams from HBM (the
"inputs larger than L2" timing rule).  The step is replayed from a CUDA graph whose launches carry
the programmatic-dependent-launch attribute.  `value` = algorithmic bytes per second over the whole job,
inputs resident in HBM, device-timed.
Other keys of the line:
  e2e                  the same step with HOST buffers (pinned x -> device, graph replay, y -> pinned host, synchronize),
                       host wall clock; next to it the eager variant (one Python -> C-ABI call per layer)
  roofline             the decode kernel against the measured HBM peak (MEASURED_PEAKS.json)
  isolated_call        one eager call, cold L2 (the reference's own timing protocol, examples/benchmark_triton.py:44-60)
  prefill              second half of the metric: M=4096 TFLOP/s at K=N=4096 (the metric's shape) and 8192 (configs[2])
  independent_launches supplementary: same 64 layers, same x for every call, 4 streams in one graph
  mx_block_scaled      supplementary: A4W4 MXFP4 / A8W8 MXFP8 dynamic layers (M=4096, K=N=8192) on the block-scaled tcgen05 kernel
  reference_triton     the UNMODIFIED reference (baseline/_ref, Triton kernels) timed on this GPU on the same shapes by
                       tools/ref_runner.py (child process): decode chained / flush, prefill -- the comparison
                       `north_star` asks for.  {"unavailable": ...} when baseline/_ref is absent.
  cpu_baseline         oracle/cpu_a16w4.c (gcc + OpenMP restatement, all host threads) on a bounded sample

With --gpus N > 1 (torchrun): every layer is ONE wide layer (N_global = 4096 * world columns, K = 4096) sharded by
output column with `ColumnShardedLinear`; a step = the 64 shard calls (same x) + ONE NCCL all-gather of the step's
outputs, captured in the same CUDA graph when NCCL allows it; weak scaling (per-GPU work fixed), value = bytes of all
ranks / max-over-ranks device time.  `mxfp4_sharded` adds BASELINE configs[4] (MXFP4 16384^2 split over the ranks).

--impl reference: the CPU arm (rank 0 only): oracle/cpu_a16w4.c on all host threads; a step = a bounded sample of the
workload (SAMPLE_LAYERS distinct layers), K steps timed as asked.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = N = 4096
GROUP = 128
LAYERS = 64
ALGO_BYTES = K * N // 2 + 2 * (K // GROUP) * N * 2 + 2 * K + 2 * N      # 8 929 280
METRIC = "A16W4 gs=128 4096^2 decode M=1 HBM GB/s"
PF_M, PF_K, PF_N = 4096, 8192, 8192
ALIAS = os.environ.get("BENCH_ALIAS", "0") == "1"               # diagnostic: all layers share one weight buffer (L2-hot)
INDEPENDENT = os.environ.get("BENCH_INDEPENDENT", "0") == "1"   # diagnostic: every layer reads the same x (no chain)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  NVML is polled in-process from a thread (a
    query takes well under a millisecond; polled every 4 ms so that the launching thread is not disturbed); nvidia-smi
    -lms is the fallback when the NVML binding is missing."""

    def __init__(self, index=0, threaded=True):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop = None, None, False
        self.max_mhz = None
        self.threaded = threaded      # False: no polling thread; the caller samples explicitly (sample()) while the GPU
                                      # executes already-enqueued work, so NVML never contends with a launch

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            # first NVML queries are slow and can hold driver locks the launch path also needs: pay for them here,
            # before the timed region starts
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            (getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons)(self.handle)
            self.nvml = pynvml
            if self.threaded:
                self.thr = threading.Thread(target=self._poll, daemon=True)
                self.thr.start()
            return self
        except Exception:
            self.nvml = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        R = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(get_reasons(self.handle))
                self.rows.append([str(mhz), str(self.max_mhz), ""] +
                                 ["Active" if (mask & R[k]) else "Not Active"
                                  for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            except Exception:
                break
            time.sleep(0.004)      # a few hundred Hz would contend with the launching thread for driver locks

    def sample(self):
        """One NVML sample from the calling thread (threaded=False mode)."""
        n = self.nvml
        if n is None:
            return
        try:
            get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            R = (getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4))
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            mask = int(get_reasons(self.handle))
            self.rows.append([str(mhz), str(self.max_mhz), ""] + ["Active" if (mask & r) else "Not Active" for r in R])
        except Exception:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.nvml is not None and self.threaded:
            try:
                self.thr.join(timeout=1)
            except Exception:
                pass
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


SAMPLE_LAYERS = 16            # CPU arm: distinct layers per step (143 MB of packed weights)


def workload_config(world):
    return {"workload": f"A16W4 gs={GROUP} decode M=1 K=N={K} fp16 (BASELINE configs[1]); step = {LAYERS} "
                        f"{'chained layers' if world == 1 else 'column shards of wide layers (N_global=%d) + one output all-gather per step' % (N * world)}",
            "l2": f"{LAYERS} distinct layers = {LAYERS * ALGO_BYTES / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
            "launch": "CUDA graph replay, programmatic dependent launch",
            "chain": ("independent layers (diagnostic)" if INDEPENDENT else "y_i feeds x_{i+1}") if world == 1 else "same x for every shard call",
            "parallelism": f"column-shard x{world}"}


def make_weights(dev, layers, seed=0, n_out=N):
    """Synthetic A16W4 layers (seeded, generated on the device). Scales are chosen so that a chained
    pass keeps |x| roughly constant: std(W) * sqrt(K) ~= 1."""
    g = torch.Generator(device=dev).manual_seed(seed)
    for _ in range(layers):
        W_q = torch.randint(0, 16, (n_out, K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
        ng = n_out * K // GROUP
        scales = (torch.rand(ng, 1, generator=g, device=dev) * 0.002 + 0.0024).half()
        zeros = (torch.rand(ng, 1, generator=g, device=dev) * 3 + 6).half()
        yield W_q, scales, zeros


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: oracle/cpu_a16w4.c (gcc + OpenMP), the only place outside tests/ and smoke() that executes oracle/ code
# ------------------------------------------------------------------------------------------------------------------
class CpuSample:
    """SAMPLE_LAYERS distinct packed A16W4 layers in host memory + the C baseline bound to them."""

    def __init__(self):
        import numpy as np
        from oracle import build_cpu
        self.np, self.C = np, build_cpu
        build_cpu.load()
        rng = np.random.default_rng(0)
        G = K // GROUP
        self.layers = []
        for _ in range(SAMPLE_LAYERS):
            Wp = rng.integers(-2 ** 31, 2 ** 31 - 1, size=(K // 8, N), dtype=np.int64).astype(np.int32)   # 8 random codes per word
            s = (rng.random((G, N), dtype=np.float32) * 0.002 + 0.0024).astype(np.float16).astype(np.float32)
            z = (-(rng.random((G, N), dtype=np.float32) * 3 + 6) * s).astype(np.float16).astype(np.float32)    # z' = -z * s
            self.layers.append((Wp, s, z))
        self.x = (rng.standard_normal((1, K), dtype=np.float32) / 10).astype(np.float16).astype(np.float32)
        self.threads = build_cpu.threads()

    def step(self):
        x = self.x
        for (Wp, s, z) in self.layers:
            y = self.C.forward(Wp, s, z, x)
            x = (y / max(1e-6, float(self.np.abs(y).max())) / 10).astype(self.np.float32)   # keep the chain in range
        return x

    def timed(self, steps, warmup):
        for _ in range(max(warmup, 1)):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / steps


def cpu_baseline_obj(steps=20, warmup=2):
    try:
        cs = CpuSample()
        sec = cs.timed(steps, warmup)
        v = SAMPLE_LAYERS * ALGO_BYTES / sec / 1e9
        return {"value": v, "unit": "GB/s", "cores": cs.threads, "kind": "port",
                "sample": f"{steps} steps x {SAMPLE_LAYERS} distinct 4096x4096 A16W4 layers (chained, M=1) through oracle/cpu_a16w4.c "
                          f"(gcc -O3 + OpenMP, fp32, group-wise factorised dequant), {sec * 1e3 / SAMPLE_LAYERS:.3f} ms/layer",
                "ms_per_layer": sec * 1e3 / SAMPLE_LAYERS}, sec
    except Exception as e:
        return {"value": 0.0, "unit": "GB/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": f"failed: {e!r}"[:200]}, float("nan")


def run_reference(args, rank, world):
    if rank != 0:
        return
    cs = CpuSample()
    sec = cs.timed(args.steps, args.warmup)
    v = SAMPLE_LAYERS * ALGO_BYTES / sec / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cs.threads, "kind": "port",
                         "sample": f"each step = {SAMPLE_LAYERS} of the {LAYERS} layers of the workload ({SAMPLE_LAYERS * ALGO_BYTES / 1e6:.0f} MB of "
                                   f"distinct packed weights, chained, M=1) through oracle/cpu_a16w4.c (gcc -O3 + OpenMP, all host "
                                   f"threads); {sec * 1e3 / SAMPLE_LAYERS:.3f} ms per 4096x4096 layer; value is a rate, so the sample size "
                                   f"does not scale it"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# secondary legs (N = 1)
# ------------------------------------------------------------------------------------------------------------------
def time_prefill(dev, pk, pn, pm=4096):
    """Second half of the metric: M=4096 TFLOP/s; flush protocol (256 MB between iterations), median of 10."""
    try:
        from gemlite_b200 import _cabi
        from gemlite_b200.helper import A16W4_HQQ_INT
        g = torch.Generator(device=dev).manual_seed(1)
        W_q = torch.randint(0, 16, (pn, pk), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
        ng = pn * pk // GROUP
        scales = (torch.rand(ng, 1, generator=g, device=dev) * 0.01 + 0.001).half()
        zeros = (torch.rand(ng, 1, generator=g, device=dev) * 15).half()
        layer = A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros)
        del W_q
        x = (torch.randn(pm, pk, generator=g, device=dev) / 10).half()
        layer(x)
        kern = _cabi.last_kernel()
        if kern == "generic_cuda_core":
            return {"available": False, "kernel": kern}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(3):
            layer(x)
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); layer(x); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        flops = 2.0 * pm * pk * pn
        peaks = measured_peaks()
        peak = peaks.get("bf16_tflops", 1590.0)
        return {"available": True, "kernel": kern, "workload": f"A16W4 gs={GROUP} prefill M={pm} K=N={pk} fp16",
                "ms": ms, "tflops": flops / ms / 1e9, "peak_tflops": peak, "frac": flops / ms / 1e9 / peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if "bf16_tflops" in peaks else "fallback 1590",
                "l2": "256 MB flush between iterations"}
    except Exception as e:  # keep the decode line alive, but say why
        return {"available": False, "error": repr(e)[:200]}


def mx_block_scaled_leg(dev, n=8192, m=4096):
    """Supplementary: the dynamic MX layers of SURVEY 8f item 1 on the block-scaled tcgen05 kernel (gl_gemm_bs.cu) --
    A4W4_MXFP_dynamic and A8W8_MXFP_dynamic, M tokens through one K = N = n layer: the whole call (activation quantiser +
    GEMM, what GemLiteLinear.__call__ costs) and the quantiser alone, events around 10 calls, 4 distinct layers > L2."""
    try:
        from gemlite_b200 import _cabi
        from gemlite_b200 import helper as H
        from gemlite_b200.quant_utils import scale_activations_mxfp4, scale_activations_mxfp8
        out = {}
        for cls, kw, quant, tag in (("A4W4_MXFP_dynamic", {}, scale_activations_mxfp4, "a4w4_mxfp4"),
                                    ("A8W8_MXFP_dynamic", {"post_scale": False}, scale_activations_mxfp8, "a8w8_mxfp8")):
            layers = []
            for i in range(4):
                lin = torch.nn.Linear(n, n, bias=False, dtype=torch.bfloat16, device=dev)
                lin.weight.requires_grad = False
                layers.append(getattr(H, cls)(device=dev, dtype=torch.bfloat16, **kw).from_linear(lin))
                del lin
            xs = [(torch.randn(m, n, device=dev) / 10).to(torch.bfloat16) for _ in range(4)]

            def timed(fn, rep=3):
                fn(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(rep):
                    fn()
                b.record(); torch.cuda.synchronize()
                return a.elapsed_time(b) / rep / 4
            call_ms = timed(lambda: [l(x) for l, x in zip(layers, xs)])
            kern = _cabi.last_kernel()
            quant_ms = timed(lambda: [quant(x) for x in xs])
            flops = 2.0 * m * n * n
            out[tag] = {"kernel": kern, "workload": f"{cls} M={m} K=N={n} bf16", "call_ms": call_ms, "quantiser_ms": quant_ms,
                        "call_tflops": flops / call_ms / 1e9, "gemm_tflops": flops / max(call_ms - quant_ms, 1e-6) / 1e9}
            del layers, xs
            torch.cuda.empty_cache()
        out["peaks"] = "profiles/tensor_peaks_r2.json: cuBLASLt fp8 e4m3 8192^3 2883-2985 TFLOP/s on this pool (no fp4 entry: no library fp4 GEMM in the image)"
        return out
    except Exception as e:
        return {"available": False, "error": repr(e)[:200]}


def isolated_call(layer, x, dev, rep=100):
    """What a caller outside a CUDA graph sees: one eager GemLiteLinear.__call__, cold L2, one event pair per call."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(rep):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); layer(x); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return {"us_min": ts[0], "us_median": ts[len(ts) // 2], "protocol": "256 MB L2 flush, event pair around one eager call "
            "(examples/benchmark_triton.py:44-60); includes the host launch path"}


def reference_triton_leg(budget_s=150.0):
    """The unmodified reference's Triton kernels on this GPU (child process; baseline/_ref)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref", "gemlite")
    if not os.path.isdir(ref_dir):
        return {"unavailable": "baseline/_ref/gemlite is absent (pip install --target baseline/_ref, DESIGN.md section 4)"}
    import tempfile
    I4 = dict(kind="pack_int", nbits=4, gs=GROUP, dtype="float16", seed=100)
    timing = [
        dict(name="decode_M1_4096", build=dict(I4, N=N, K=K), M=1, n_layers=LAYERS, chained=True, rep=100),
        dict(name="prefill_M4096_4096", build=dict(I4, N=4096, K=4096), M=4096, n_layers=8, rep=20),
        dict(name="prefill_M4096_8192", build=dict(I4, N=8192, K=8192), M=4096, n_layers=4, rep=20),
    ]
    tmp = tempfile.mkdtemp(prefix="bench_ref_")
    job, out = os.path.join(tmp, "job.pt"), os.path.join(tmp, "out.pt")
    torch.save({"cases": [], "timing": timing, "autotune": "fast", "budget_s": budget_s}, job)
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    t0 = time.time()
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_runner.py"), job, out], env=env, cwd=tmp,
                       timeout=budget_s + 120, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except subprocess.TimeoutExpired:
        pass
    if not os.path.exists(out):
        return {"unavailable": "reference child produced no result (see tools/ref_runner.py)"}
    r = torch.load(out, weights_only=False)
    t = r["timing"]
    res = {"version": r.get("version"), "triton": r.get("triton"), "autotune": "fast (the reference's default mode; no b200.json ships)",
           "seconds": time.time() - t0, "protocols": "graph: distinct layers > L2 in one CUDA graph, ms per call; flush: eager call after a "
           "256 MB L2 flush (the reference's own protocol)", "raw": t}
    d = t.get("decode_M1_4096", {})
    if "graph_chained_ms_per_layer" in d:
        us = d["graph_chained_ms_per_layer"] * 1e3
        res["decode"] = {"us_per_layer_chained_graph": us, "value": ALGO_BYTES / us / 1e3, "unit": "GB/s",
                         "us_flush_median": d.get("flush_median_ms", 0) * 1e3}
    for key, (pk, pn) in (("prefill_M4096_4096", (4096, 4096)), ("prefill_M4096_8192", (8192, 8192))):
        d = t.get(key, {})
        if "graph_ms_per_layer" in d:
            res[key] = {"ms_graph": d["graph_ms_per_layer"], "ms_flush_median": d.get("flush_median_ms"),
                        "tflops": 2.0 * 4096 * pk * pn / d["graph_ms_per_layer"] / 1e9}
    return res


def capture_independent(layers, x, n_streams, dev):
    """CUDA graph of one call per layer, every call reading the SAME x (no y -> x dependency), spread round-robin
    over n_streams streams (fork / join inside the capture).  Returns (graph, outputs)."""
    main = torch.cuda.Stream(device=dev)
    side = [main] + [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
    keep = []

    def body():
        fork = torch.cuda.Event()
        fork.record(main)
        for s in side[1:]:
            s.wait_event(fork)
        for i, layer in enumerate(layers):
            with torch.cuda.stream(side[i % n_streams]):
                keep.append(layer(x))
        for s in side[1:]:
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)

    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(main):
        body()                       # eager warm-up on the same streams
        torch.cuda.synchronize()
        keep.clear()
        with torch.cuda.graph(graph, stream=main):
            body()
    return graph, keep


def time_graph_us_per_layer(graph, steps, n_layers):
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (steps * n_layers)


def independent_leg(layers, x, dev, peak, n_streams=4, steps=30):
    """Supplementary number (never the headline): the same layers without the y -> x chain, on n_streams streams."""
    try:
        graph, keep = capture_independent(layers, x, n_streams, dev)
        us = time_graph_us_per_layer(graph, steps, len(layers))
        gbs = ALGO_BYTES / (us * 1e-6) / 1e9
        return {"streams": n_streams, "us_per_launch": us, "value": gbs, "unit": "GB/s", "frac_of_peak": gbs / peak,
                "note": "same 64 layers, every call reads the same x (independent launches), CUDA-graph replay; "
                        "supplementary, not the headline workload"}
    except Exception as e:      # supplementary only: never lose the bench line over it
        return {"error": f"{type(e).__name__}: {str(e).splitlines()[0][:160] if str(e) else ''}"}


def mxfp4_sharded_leg(dev, rank, world, steps=10):
    """BASELINE configs[4]: MXFP4 gs=32 K=N=16384 column-sharded over the ranks (strong scaling: total work fixed),
    kernel time and kernel + all-gather time at M=1 and M=4096, device-timed, max over ranks."""
    import torch.distributed as dist
    from gemlite_b200.helper import A16W4_MXFP
    from gemlite_b200.sharding import ColumnShardedLinear
    NK = 16384
    out = {"workload": f"A16W4 MXFP4 gs=32 K=N={NK} bf16, {world} column shards", "per_M": {}}
    try:
        g = torch.Generator(device=dev).manual_seed(123)
        n_loc = NK // world
        # the shard of a layer whose codes / scales are random: generated shard-locally with a rank-dependent seed
        W_q = torch.randint(0, 16, (n_loc, NK), generator=torch.Generator(device=dev).manual_seed(1000 + rank), device=dev, dtype=torch.uint8)
        sc = torch.randint(118, 126, (n_loc, NK // 32), generator=g, device=dev, dtype=torch.uint8).view(torch.float8_e8m0fnu)
        local = A16W4_MXFP(device=dev, dtype=torch.bfloat16).from_weights(W_q, sc)
        del W_q
        for M in (1, 4096):
            x = (torch.randn(M, NK, device=dev, generator=g) / 10).bfloat16()
            y = local(x)
            gathered = torch.empty(world, M, n_loc, dtype=torch.bfloat16, device=dev)

            def timed(with_gather):
                for _ in range(3):
                    yy = local(x)
                    if with_gather:
                        dist.all_gather_into_tensor(gathered.view(-1), yy.view(-1))
                torch.cuda.synchronize(); dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(steps):
                    yy = local(x)
                    if with_gather:
                        dist.all_gather_into_tensor(gathered.view(-1), yy.view(-1))
                b.record(); torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return t.item()

            k_ms, kg_ms = timed(False), timed(True)
            byts = (NK * NK // 2 + NK * NK // 32) / world
            out["per_M"][str(M)] = {"kernel_ms": k_ms, "kernel_plus_gather_ms": kg_ms,
                                    "per_gpu_weight_GBps": byts / k_ms / 1e6,
                                    "aggregate_tflops": 2.0 * M * NK * NK / kg_ms / 1e9, "gather_bytes_per_rank": M * n_loc * 2}
            del x, y, gathered
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {str(e).splitlines()[0][:200] if str(e) else ''}"
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from gemlite_b200 import _cabi
    from gemlite_b200.helper import A16W4_HQQ_INT
    from gemlite_b200.sharding import ColumnShardedLinear

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    layers, shards = [], []
    if world == 1:
        for (W_q, scales, zeros) in make_weights(dev, 1 if ALIAS else LAYERS, seed=0):
            layers.append(A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros))
            del W_q
        if ALIAS:
            layers = layers * LAYERS
    else:
        # ONE wide layer per step position (same seed on every rank => the same global weights), sharded by output column
        for (W_q, scales, zeros) in make_weights(dev, LAYERS, seed=0, n_out=N * world):
            full = A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros)
            del W_q, scales, zeros
            shards.append(ColumnShardedLinear(full, rank, world))
            layers.append(shards[-1].local)
            del full
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    x0 = (torch.randn(1, K, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) / 10).half()
    x_in = x0.clone()          # static input buffer of the captured step (the e2e leg copies host data into it)

    outs = torch.empty(LAYERS, 1, N, dtype=torch.float16, device=dev)
    gathered = torch.empty(world * LAYERS, 1, N, dtype=torch.float16, device=dev) if world > 1 else None

    def step_body(with_gather):
        x = x_in
        if world > 1:
            ys = [layer(x) for layer in layers]
            torch.stack(ys, out=outs)            # one concatenation at the end: no foreign kernel between decode launches
            if with_gather:
                dist.all_gather_into_tensor(gathered.view(-1), outs.view(-1))
            return outs
        for layer in layers:
            y = layer(x)
            if not INDEPENDENT:
                x = y
        return x

    # warm up eagerly (module load, smem opt-in, NCCL communicator), then capture the step in a CUDA graph
    for _ in range(2):
        step_body(True)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)
    gather_in_graph = False

    def capture(with_gather):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            step_body(with_gather)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=stream):
                r = step_body(with_gather)
        return g, r

    if world > 1:
        try:
            graph, result = capture(True)
            gather_in_graph = True
        except Exception:
            torch.cuda.synchronize()
            graph, result = capture(False)
    else:
        graph, result = capture(False)

    def nccl_step():
        graph.replay()
        if world > 1 and not gather_in_graph:
            dist.all_gather_into_tensor(gathered.view(-1), outs.view(-1))

    # B200-native exchange for M = 1: the decode kernel's epilogue stores its 32 outputs into every rank's
    # symmetric output buffer over NVLink; one barrier pair per step replaces stack + all-gather.  NCCL stays the baseline.
    fused = {"available": False}
    fgraph = None
    if world > 1 and os.environ.get("BENCH_NO_FUSED_GATHER", "0") != "1":
        try:
            from gemlite_b200.sharding import SymmetricOutputs
            sym, sym_err = None, None
            try:
                sym = SymmetricOutputs(LAYERS, N * world, torch.float16, dev)
            except Exception as e:
                sym_err = e
            # all ranks take the fused path or none does: a rank that failed alone would leave the others in a barrier
            ok = torch.tensor([0 if sym is None else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                raise sym_err if sym_err is not None else RuntimeError("symmetric buffers unavailable on another rank")

            def fused_body():
                sym.barrier()                       # every rank has finished reading the previous step's rows
                for i, sh in enumerate(shards):
                    sh.forward_scatter(x_in, sym, i)
                sym.barrier()                       # every rank's columns have landed everywhere

            fused_body()
            torch.cuda.synchronize()
            nccl_step()
            torch.cuda.synchronize()
            same = bool(torch.equal(sym.buf.view(LAYERS, world, N), gathered.view(world, LAYERS, N).transpose(0, 1)))
            fgraph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(stream):
                fused_body()
                torch.cuda.synchronize()
                with torch.cuda.graph(fgraph, stream=stream):
                    fused_body()
            fused = {"available": True, "equals_nccl_gather": same,
                     "how": "decode epilogue stores to every rank's symmetric buffer (glb_matmul_bcast) + 2 symmetric-memory barriers per step, in one CUDA graph"}
        except Exception as e:      # keep the NCCL line alive, but say why
            fused = {"available": False, "error": f"{type(e).__name__}: {str(e).splitlines()[0][:200] if str(e) else ''}"}
            fgraph = None
            torch.cuda.synchronize()

    def one_step():
        if fgraph is not None:
            fgraph.replay()
        else:
            nccl_step()

    # NCCL connects channels lazily (also for graph-captured collectives): with N > 1 the first replays are milliseconds
    # slower than the steady state, so more untimed steps than asked are run (reported as `warmup`)
    n_warm = max(args.warmup, 3) if world == 1 else max(args.warmup, 40)
    for _ in range(n_warm):
        one_step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    # NVML is initialised BEFORE the barrier: done after it (as in the first 2- and 4-GPU runs of this round) the ranks left
    # the barrier aligned and then spent different milliseconds in nvmlInit, and the early ranks waited for the late one in
    # the first step's device-side barrier -- one ~1 ms step out of 20 (profiles/bench_r2_n4_visit21.json)
    with ClockSampler(local_rank, threaded=False) as clocks:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        marks[0].record()
        for i in range(args.steps):
            one_step()
            marks[i + 1].record()
        ev1.record()
        # the K steps are enqueued (graph launches are asynchronous) and executing: sample the clocks now, from this
        # thread, instead of from a polling thread that can hold driver locks a launch needs
        for _ in range(3):
            clocks.sample()
        torch.cuda.synchronize()
        clocks.sample()
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_step = ms_total / args.steps
    value = world * LAYERS * ALGO_BYTES / (ms_step * 1e-3) / 1e9

    nccl_ms = None
    if world > 1 and fgraph is not None:            # the NCCL variant of the same step, same timing rules (reported next to it)
        for _ in range(n_warm):
            nccl_step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            nccl_step()
        b.record(); torch.cuda.synchronize(); dist.barrier()
        tt = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        nccl_ms = tt.item()

    # kernel-only time inside the step (the decode launches without the gather), for the roofline object
    if world > 1 and gather_in_graph:
        kgraph, _ = capture(False)
    else:
        kgraph = graph
    for _ in range(3):
        kgraph.replay()
    torch.cuda.synchronize()
    evk0, evk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evk0.record()
    for _ in range(args.steps):
        kgraph.replay()
    evk1.record()
    torch.cuda.synchronize()
    us_per_launch = evk0.elapsed_time(evk1) * 1e3 / (args.steps * LAYERS)

    # end to end with host buffers: H2D of the step input + the step + D2H of the result inside the timed region
    xh = x0.cpu().pin_memory()
    yh = torch.empty(1, N, dtype=torch.float16).pin_memory()
    y_static = result if world == 1 else None

    def e2e_graph_step():
        x_in.copy_(xh, non_blocking=True)
        one_step()
        if world > 1:
            yh.copy_((sym.buf[-1, -N:] if fgraph is not None else gathered[-1]).view(1, N), non_blocking=True)
        else:
            yh.copy_(y_static, non_blocking=True)
        torch.cuda.synchronize()

    def e2e_eager_step():
        x_in.copy_(xh, non_blocking=True)
        x = x_in
        for layer in layers:
            x = layer(x)
        yh.copy_(x, non_blocking=True)
        torch.cuda.synchronize()

    def time_e2e(fn):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    e2e_s = time_e2e(e2e_graph_step)
    l0 = _cabi.launch_count()
    eager_s = time_e2e(e2e_eager_step) if world == 1 else float("nan")
    eager_launches = _cabi.launch_count() - l0
    e2e_val = world * LAYERS * ALGO_BYTES / e2e_s / 1e9
    eager_val = LAYERS * ALGO_BYTES / eager_s / 1e9 if world == 1 else None
    skip = os.environ.get("BENCH_SKIP_EXTRAS", "0") == "1"     # diagnostic sweeps only
    mx_sharded = mxfp4_sharded_leg(dev, rank, world) if (world > 1 and not skip) else None

    if rank != 0:
        return
    peaks = measured_peaks()
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = ALGO_BYTES / (us_per_launch * 1e-6) / 1e9
    traffic, traffic_src = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "decode_traffic.json")))
        traffic = prof.get("dram_bytes_per_launch")
        traffic_src = prof.get("source", "profiles/decode_traffic.json (ncu --set full capture, not re-measured by this run)")
    except Exception:
        pass
    cpu_obj = None
    if world == 1 and not skip:
        cpu_obj, _ = cpu_baseline_obj()
    prefill = None
    if world == 1 and not skip:
        prefill = time_prefill(dev, PF_K, PF_N)
        prefill["at_4096"] = time_prefill(dev, 4096, 4096)
    independent = independent_leg(layers, x0, dev, peak) if (world == 1 and not skip and not ALIAS) else None
    isolated = isolated_call(layers[0], x0, dev) if (world == 1 and not skip) else None
    mx_bs = mx_block_scaled_leg(dev) if (world == 1 and not skip) else None
    ref_triton = reference_triton_leg() if (world == 1 and not skip and os.environ.get("BENCH_SKIP_REFERENCE_TRITON", "0") != "1") else None
    cfg = workload_config(world)
    handoff = {"handoff": "griddepcontrol.wait (programmatic dependent launch); a device-side arrival counter was measured slower (DESIGN.md 3.3)"}
    if world > 1:
        handoff["gather"] = ("fused: " + fused["how"]) if fgraph is not None else \
            ("NCCL all-gather captured in the step's CUDA graph" if gather_in_graph else "NCCL all-gather issued eagerly after the graph replay")
        handoff["fused_gather"] = fused
        handoff["nccl_baseline"] = None if nccl_ms is None else {
            "ms_per_step": nccl_ms, "value": world * LAYERS * ALGO_BYTES / (nccl_ms * 1e-3) / 1e9, "unit": "GB/s",
            "gather": "captured in the graph" if gather_in_graph else "eager"}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_step, "ms_per_step_rank0_min_median_max": [per_step[0], per_step[len(per_step) // 2], per_step[-1]], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic", "config": cfg,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": "gemv_w4_m1_kernel",
                     "us_per_launch": us_per_launch, "algorithmic_bytes_per_launch": ALGO_BYTES,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"},
        "cpu_baseline": cpu_obj,
        "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": 2 * K, "d2h_bytes_per_step": 2 * N,
                "ms_per_step": e2e_s * 1e3,
                "path": "pinned host x -> device, CUDA-graph replay of the 64 GemLiteLinear.__call__s, y -> pinned host, synchronize",
                "eager_value": eager_val, "eager_ms_per_step": (eager_s * 1e3 if world == 1 else None),
                "eager_path": "same, one eager Python -> C-ABI call per layer"},
        "gpu_launches": args.steps * LAYERS, "gpu_launches_e2e_eager": int(eager_launches), "decode_handoff": handoff,
        "clocks": clocks.summary(),
        "prefill": prefill,
        "isolated_call": isolated,
        "independent_launches": independent,
        "reference_triton": ref_triton,
        "mxfp4_sharded": mx_sharded,
        "mx_block_scaled": mx_bs,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
