"""Diagnostic: how fast does the decode kernel stream when consecutive launches do NOT depend on each other?

bench.py's step chains 64 layers (y_i feeds x_{i+1}); its per-layer time is bounded by the dependent-launch hand-off
plus the post-wait critical path, not by HBM.  This tool replays the same 64 distinct layers with the
same x for every layer, spread round-robin over S streams inside one CUDA graph (fork / join), so S independent
dependent-launch chains overlap on the device.  S = 1 is the single-stream independent case (every launch still
executes griddepcontrol.wait, so it matches the chained time); S > 1 shows what the kernel itself sustains once the
launch chain is off the critical path.

  python tools/decode_streams.py [--streams 1,2,3,4] [--steps 50]

Prints one line per S and appends them to gpurun_out/decode_streams.txt.  Not a bench value: the workload differs
from bench.py's (independent instead of chained layers)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (make_weights, sizes)


def build(dev):
    from gemlite_b200.helper import A16W4_HQQ_INT
    layers = []
    for (W_q, scales, zeros) in bench.make_weights(dev, bench.LAYERS, seed=0):
        layers.append(A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, scales, zeros))
        del W_q
    torch.cuda.synchronize()
    return layers


def capture(layers, x, S, chained, dev):
    if not chained:
        return bench.capture_independent(layers, x, S, dev)
    stream = torch.cuda.Stream(device=dev)
    keep = []

    def body():
        y = x
        for layer in layers:
            y = layer(y)
        keep.append(y)

    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        body()
        torch.cuda.synchronize()
        keep.clear()
        with torch.cuda.graph(graph, stream=stream):
            body()
    return graph, keep


def time_graph(graph, steps):
    return bench.time_graph_us_per_layer(graph, steps, bench.LAYERS)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", default="1,2,3,4")
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    layers = build(dev)
    x = (torch.randn(1, bench.K, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) / 10).half()
    ref = [layer(x) for layer in layers]
    torch.cuda.synchronize()
    lines = []
    cases = [("chained", 1)] + [("independent", int(s)) for s in args.streams.split(",")]
    for kind, S in cases:
        try:
            graph, keep = capture(layers, x, S, kind == "chained", dev)
            us = time_graph(graph, args.steps)
            ok = ""
            if kind == "independent":
                same = all(torch.equal(a, b) for a, b in zip(keep, ref))
                ok = " outputs_identical_to_eager=%s" % same
            gbs = bench.ALGO_BYTES / (us * 1e-6) / 1e9
            lines.append("variant=%s %s streams=%d us/layer=%.3f GB/s=%.0f%s" % (
                os.environ.get("GLB_GEMV_VARIANT", "default"), kind, S, us, gbs, ok))
        except Exception as e:  # keep going: this is a one-shot diagnostic
            lines.append("variant=%s %s streams=%d FAILED %s: %s" % (
                os.environ.get("GLB_GEMV_VARIANT", "default"), kind, S, type(e).__name__, str(e).splitlines()[0][:200]))
        print(lines[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "decode_streams.txt"), "a") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
