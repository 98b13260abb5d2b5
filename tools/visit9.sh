#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== full suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v9.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v9.log | cut -c1-330 | head -30
echo "== launch list of the bench step (our kernels only)"
BENCH_SKIP_EXTRAS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_|gemm_|generic_kernel|act_quant|pack_cols|expand_act" -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r2.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r2.csv')) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ik][:90], [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for n, (c, t) in agg.items():
    print(f"{c:5d} launches {t/1e3:10.2f} us total {100*t/tot:5.1f}%  avg {t/c/1e3:7.2f} us  {n}")
PY
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
