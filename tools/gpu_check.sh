#!/bin/bash
# One GPU-box visit that re-validates everything the tree claims: parity suite, smoke, bench (both arms), the dynamic-MX
# breakdown, ours-vs-reference timings of the formats added in round 2, and ncu captures of the top kernels.
# Everything is written under gpurun_out/ (merged back by gpurun).  ~12 minutes of box time.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --tb=short > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/pytest_gpu_full.log | cut -c1-300 | head -60
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 800 gpurun_out/bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
echo "== dynamic MX layers: quantiser / GEMM breakdown"; timeout 300 python tools/run_mx.py > gpurun_out/mx_breakdown.txt 2>&1; cut -c1-260 gpurun_out/mx_breakdown.txt
echo "== ours vs reference (formats of round 2)"
timeout 900 python tools/bench_vs_reference.py --only a16w8,a16w2,mxdyn --budget 500 --out gpurun_out/vs_reference_timing.json 2>&1 | grep -v "^\[ours\]\|^\[ref\]" | tail -60
echo "== ncu launch list (bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    env BENCH_SKIP_REFERENCE_TRITON=1 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: decode, prefill, block-scaled"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_m1 -s 140 -c 1 -o gpurun_out/decode_full -f \
    env BENCH_SKIP_EXTRAS=1 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a16w4_tc2 -c 1 -o gpurun_out/prefill_full -f \
    python tools/run_prefill.py --quick > gpurun_out/ncu_prefill.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_mx_bs -s 2 -c 1 -o gpurun_out/bs_mxfp4 -f \
    python tools/run_bs_once.py A4W4_MXFP_dynamic 8192 8192 4096 > gpurun_out/ncu_bs.log 2>&1
ls -la gpurun_out | tail -20
