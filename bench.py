#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at N=1 (BASELINE.json configs[1]): A16W4 group-size-128 decode, M=1, K=N=4096, fp16.
One *step* = one decode pass through LAYERS=64 distinct layers of that shape, chained
(y of layer i is x of layer i+1), i.e. 572 MB of packed weights + metadata per step -- 4.5x the
126 MB L2, so every layer streams from HBM (this is the "inputs larger than L2" timing rule).
`value` is algorithmic bytes moved per second over the whole job, inputs resident in HBM; the step is
replayed from a CUDA graph whose launches carry the programmatic-dependent-launch attribute.
`e2e` is the same step end to end with HOST buffers: the step input is copied from pinned host memory, the
64 `GemLiteLinear.__call__`s run (replayed from the CUDA graph a serving stack captures around its decode
step -- torch.cuda.CUDAGraph, the public PyTorch facility) and the output is copied back to pinned host memory,
all inside the timed region (host wall clock incl. the final synchronize).  The same loop with plain eager
calls (one Python -> C-ABI call per layer) is reported next to it as `e2e.eager_value`.
`prefill` reports the second half of the metric (M=4096 TFLOP/s) on configs[2].
`independent_launches` (N=1, supplementary, never the headline): the same 64 layers with the same x for every call (no
y -> x chain) on 4 streams inside one CUDA graph -- what the kernel streams once the dependent-launch hand-off is off
the critical path.

With --gpus N > 1 (torchrun): weak scaling -- every rank owns its own column shard (4096 columns)
of LAYERS wide layers (global N = 4096*world), x replicated, and the step ends with ONE NCCL
all-gather of the step's outputs; value = bytes of all ranks / max-over-ranks time.

--impl reference: times the CPU restatement of the reference path (oracle port: torch dequant +
matmul, all host threads) on a bounded sample of the same workload, rank 0 only.
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

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop = None, None, False
        self.max_mhz = None

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

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.nvml is not None:
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


def make_weights(dev, layers, seed=0):
    """Synthetic A16W4 layers (seeded, generated on the device). Scales are chosen so that a chained
    pass keeps |x| roughly constant: std(W) * sqrt(K) ~= 1."""
    g = torch.Generator(device=dev).manual_seed(seed)
    out = []
    for _ in range(layers):
        W_q = torch.randint(0, 16, (N, K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
        ng = N * K // GROUP
        scales = (torch.rand(ng, 1, generator=g, device=dev) * 0.002 + 0.0024).half()
        zeros = (torch.rand(ng, 1, generator=g, device=dev) * 3 + 6).half()
        out.append((W_q, scales, zeros))
    return out


def cpu_reference_sample(threads, seconds=10.0, max_calls=200):
    """Oracle port (torch dequant + matmul, fp32) of one 4096^2 A16W4 layer on the host cores."""
    from oracle import gemlite_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    W_q = torch.randint(0, 16, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    scales = (torch.rand(N * K // GROUP, 1, generator=g) * 0.002 + 0.0024).half()
    zeros = (torch.rand(N * K // GROUP, 1, generator=g) * 3 + 6).half()
    p = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=GROUP, in_features=K, out_features=N)
    Wp = p.W_q.contiguous()
    x = (torch.randn(1, K, generator=g) / 10).half()
    O.cpu_dequant_matmul(Wp, p.scales, p.zeros, x, GROUP)
    t0 = time.perf_counter()
    calls = 0
    while calls < max_calls and (time.perf_counter() - t0) < seconds:
        O.cpu_dequant_matmul(Wp, p.scales, p.zeros, x, GROUP)
        calls += 1
    dt = time.perf_counter() - t0
    return ALGO_BYTES * calls / dt / 1e9, calls, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # a "step" = LAYERS layer calls; the CPU arm runs a bounded sample of it (calls are identical work)
    per_step_calls = 4
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample(threads, seconds=1.0, max_calls=per_step_calls)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, calls, dt = cpu_reference_sample(threads, seconds=30.0, max_calls=per_step_calls)
        vals.append((v, dt / calls))
    v = sum(a for a, _ in vals) / len(vals)
    ms_call = 1e3 * sum(b for _, b in vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_call * LAYERS, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"A16W4 gs={GROUP} decode M=1 K=N={K}, {LAYERS} layers/step "
                               f"(CPU arm: {per_step_calls} layer calls sampled per step, extrapolated)"},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step_calls} calls/step x {args.steps} steps of torch unpack+dequant+matmul (fp32) "
                                   f"on one 4096x4096 A16W4 layer, {ms_call:.1f} ms/call"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_prefill(dev):
    """Second half of the metric: M=4096 on K=N=8192 (configs[2]); returns dict or None if not available."""
    try:
        from gemlite_b200 import _cabi
        from gemlite_b200.helper import A16W4_HQQ_INT
        g = torch.Generator(device=dev).manual_seed(1)
        W_q = torch.randint(0, 16, (PF_N, PF_K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
        ng = PF_N * PF_K // GROUP
        scales = (torch.rand(ng, 1, generator=g, device=dev) * 0.01 + 0.001).half()
        zeros = (torch.rand(ng, 1, generator=g, device=dev) * 15).half()
        layer = A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros)
        del W_q
        x = (torch.randn(PF_M, PF_K, generator=g, device=dev) / 10).half()
        y = layer(x)
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
        flops = 2.0 * PF_M * PF_K * PF_N
        peaks = measured_peaks()
        peak = peaks.get("bf16_tflops", 1590.0)
        return {"available": True, "kernel": kern, "workload": f"A16W4 gs={GROUP} prefill M={PF_M} K=N={PF_K} fp16",
                "ms": ms, "tflops": flops / ms / 1e9, "peak_tflops": peak,
                "frac": flops / ms / 1e9 / peak, "l2": "256 MB flush between iterations"}
    except Exception as e:  # keep the decode line alive, but say why
        return {"available": False, "error": repr(e)[:200]}


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
    """Supplementary number (never the headline): the same layers without the y -> x chain, on n_streams streams.
    Shows what the decode kernel streams once the dependent-launch hand-off is off the critical path."""
    try:
        graph, keep = capture_independent(layers, x, n_streams, dev)
        us = time_graph_us_per_layer(graph, steps, len(layers))
        gbs = ALGO_BYTES / (us * 1e-6) / 1e9
        return {"streams": n_streams, "us_per_launch": us, "value": gbs, "unit": "GB/s", "frac_of_peak": gbs / peak,
                "note": "same 64 layers, every call reads the same x (independent launches), CUDA-graph replay; "
                        "supplementary, not the headline workload"}
    except Exception as e:      # supplementary only: never lose the bench line over it
        return {"error": f"{type(e).__name__}: {str(e).splitlines()[0][:160] if str(e) else ''}"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from gemlite_b200 import _cabi
    from gemlite_b200.helper import A16W4_HQQ_INT

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    layers = []
    for (W_q, scales, zeros) in make_weights(dev, 1 if ALIAS else LAYERS, seed=rank):
        layers.append(A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros))
        del W_q
    if ALIAS:
        layers = layers * LAYERS
    torch.cuda.synchronize()
    x0 = (torch.randn(1, K, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) / 10).half()
    x_in = x0.clone()          # static input buffer of the captured step (the e2e leg copies host data into it)

    # device-resident step: chained decode pass (N=1) or independent shards of wide layers + one all-gather
    outs = torch.empty(LAYERS, 1, N, dtype=torch.float16, device=dev)
    gathered = torch.empty(world * LAYERS, 1, N, dtype=torch.float16, device=dev) if world > 1 else None

    def step_body():
        x = x_in
        if world > 1:
            # shard outputs are collected for the single gather with ONE concatenation at the end of the step (a copy
            # per layer would put a non-PDL kernel between every two decode launches)
            ys = [layer(x) for layer in layers]
            torch.stack(ys, out=outs)
            return outs
        for layer in layers:
            y = layer(x)
            if not INDEPENDENT:
                x = y
        return x

    # warm up eagerly (module load, smem opt-in), then capture the step in a CUDA graph
    for _ in range(2):
        step_body()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        step_body()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            result = step_body()

    def one_step():
        graph.replay()
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), outs.view(-1))

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            one_step()
        ev1.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_step = ms_total / args.steps
    value = world * LAYERS * ALGO_BYTES / (ms_step * 1e-3) / 1e9

    # kernel-only time inside the step (graph replay without the gather), for the roofline object
    evk0, evk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evk0.record()
    for _ in range(args.steps):
        graph.replay()
    evk1.record()
    torch.cuda.synchronize()
    us_per_launch = evk0.elapsed_time(evk1) * 1e3 / (args.steps * LAYERS)

    # end to end with host buffers: H2D of the step input + the step + D2H of the result inside the timed region
    xh = x0.cpu().pin_memory()
    yh = torch.empty(1, N, dtype=torch.float16).pin_memory()
    y_static = result if world == 1 else None

    def e2e_graph_step():
        x_in.copy_(xh, non_blocking=True)
        graph.replay()
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), outs.view(-1))
            yh.copy_(gathered[-1], non_blocking=True)
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

    if rank != 0:
        return
    peaks = measured_peaks()
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = ALGO_BYTES / (us_per_launch * 1e-6) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "decode_traffic.json")))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    skip = os.environ.get("BENCH_SKIP_EXTRAS", "0") == "1"     # diagnostic sweeps only
    # the CPU baseline is reported at N = 1 only (rank 0); multi-GPU runs do not repeat it
    cpu_v, cpu_calls, cpu_dt = (0.0, 0, 0.0) if (skip or world > 1) else cpu_reference_sample(os.cpu_count() or 1, seconds=12.0, max_calls=100)
    prefill = time_prefill(dev) if (world == 1 and not skip) else None
    independent = independent_leg(layers, x0, dev, peak) if (world == 1 and not skip and not ALIAS) else None
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": f"A16W4 gs={GROUP} decode M=1 K=N={K} fp16 (BASELINE configs[1]); step = {LAYERS} "
                               f"{'chained layers' if world == 1 else 'column shards (N_global=%d) + one NCCL all-gather' % (N * world)}",
                   "l2": f"{LAYERS} distinct layers = {LAYERS * ALGO_BYTES / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                   "launch": "CUDA graph replay, programmatic dependent launch", "chain": "independent layers (diagnostic)" if INDEPENDENT else "y_i feeds x_{i+1}", "parallelism": f"column-shard x{world}"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "gemv_w4_tma_kernel", "us_per_launch": us_per_launch,
                     "algorithmic_bytes_per_launch": ALGO_BYTES,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"},
        "cpu_baseline": ({"value": cpu_v, "unit": "GB/s", "cores": os.cpu_count() or 1, "kind": "port",
                          "sample": f"{cpu_calls} calls of torch unpack+dequant+matmul (fp32) on one 4096x4096 A16W4 layer in {cpu_dt:.1f} s"}
                         if world == 1 else None),
        "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": 2 * K, "d2h_bytes_per_step": 2 * N,
                "ms_per_step": e2e_s * 1e3,
                "path": "pinned host x -> device, CUDA-graph replay of the 64 GemLiteLinear.__call__s, y -> pinned host, synchronize",
                "eager_value": eager_val, "eager_ms_per_step": (eager_s * 1e3 if world == 1 else None),
                "eager_path": "same, one eager Python -> C-ABI call per layer"},
        "gpu_launches": args.steps * LAYERS, "gpu_launches_e2e_eager": int(eager_launches),
        "clocks": clocks.summary(),
        "prefill": prefill,
        "independent_launches": independent,
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
        if args.steps > 5:
            args.steps = 5      # bounded CPU sample
        args.warmup = min(args.warmup, 1)
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
