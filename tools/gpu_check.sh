#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list and one full ncu capture of the decode kernel.
# Everything is written under gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest" ; timeout 1200 python -m pytest tests -m gpu -q --timeout=300 --tb=short > gpurun_out/pytest_gpu_full.log 2>&1; grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/pytest_gpu_full.log | cut -c1-300 | head -80 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== bench nopdl (skipped)"
echo "== ncu launch list (our kernels)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_w4|gemm_a16w4|generic_kernel|act_quant|pack_cols" -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full (decode kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_w4 -s 130 -c 3 -o gpurun_out/decode_full -f \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full (prefill kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a16w4_tc2 -c 1 -o gpurun_out/prefill_full -f \
    python tools/run_prefill.py --quick > gpurun_out/ncu_prefill.log 2>&1
echo "== prefill sweep"; timeout 300 python tools/run_prefill.py 2>&1 | tail -6 | tee gpurun_out/prefill_sweep.log
echo "== config table"; timeout 600 python tools/bench_configs.py > gpurun_out/configs_final.txt 2>&1; tail -45 gpurun_out/configs_final.txt | cut -c1-200
ls -la gpurun_out | tail -20
