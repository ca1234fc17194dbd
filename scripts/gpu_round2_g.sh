#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'dense_|fused_' -s 34 -c 36 --csv --log-file gpurun_out/launches_dense_nocc_r02.csv python bench_configs.py --only c3 --reps 1 --fused-only > gpurun_out/g_ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_dense_nocc_r02.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines))
agg={}
for r in rows:
    k=(int(r['ID']), r['Kernel Name'].replace('b2::','')[:28])
    agg.setdefault(k,{})[r['Metric Name'].split('.')[0][-10:]]=r['Metric Value']
for k in sorted(agg): print(k, agg[k])
PY
