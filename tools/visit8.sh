#!/bin/bash
# round 2: MX expansion path tests, final ncu captures for profiles/, full suite
set -u
mkdir -p gpurun_out
echo "== MX dynamic tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "mx_dynamic or fixtures or mxfp" > gpurun_out/pytest_v8_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v8_new.log | cut -c1-330 | head -30
echo "== launch list of the bench step (decode only)"
BENCH_SKIP_EXTRAS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r2.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r2.csv')) if len(r) > 5]
hdr = rows[0]
try:
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[ik][:70]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    for n, (c, t) in agg.items():
        print(f"{c:5d} launches {t/1e3:10.2f} us total {100*t/tot:5.1f}%  {n}")
except Exception as e:
    print("parse failed", e, hdr[:8])
PY
echo "== ncu full captures"
cap() { name=$1; kre=$2; skip=$3; shift 3; timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c 1 -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log | cut -c1-100; }
BENCH_SKIP_EXTRAS=1 cap decode_v7_r2 gemv_w4_m1 140 python bench.py --steps 2 --warmup 1
cap prefill_tc2_r2 gemm_a16w4_tc2 2 python tools/run_prefill.py --quick
cap mid_w4_r2 gemm_a16w4_mid 2 python tools/run_mid.py 64 w4
cap mid_a8_r2 gemm_a8w8_mid 2 python tools/run_mid.py 64 a8
ls -la gpurun_out/*_r2.ncu-rep | awk '{print $5, $9}'
echo "== full suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v8.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v8.log | cut -c1-330 | head -30
