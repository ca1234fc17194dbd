#!/bin/bash
# round-2 session-2 batch 7: group-by / dictionary_encode with a large_utf8 key -- timing, per-kernel launch list, bench line
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/y_host_test.log 2>&1; echo "host test rc=$?"; tail -3 gpurun_out/y_host_test.log | cut -c1-300
timeout 600 python bench_configs.py --rows 40000000 --only c3u > gpurun_out/y_c3u_small.jsonl 2> gpurun_out/y_c3u_small.err; echo "small rc=$?"
cat gpurun_out/y_c3u_small.jsonl | cut -c1-400; tail -5 gpurun_out/y_c3u_small.err
timeout 900 python bench_configs.py --only c3u > gpurun_out/y_c3u.jsonl 2> gpurun_out/y_c3u.err; echo "full rc=$?"
cat gpurun_out/y_c3u.jsonl | cut -c1-400; tail -5 gpurun_out/y_c3u.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3u_r02.csv \
    python bench_configs.py --rows 400000000 --reps 1 --only c3u > gpurun_out/y_c3u_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_c3u_r02.csv', errors='ignore')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r]
h = rows[hi[0]]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi[0] + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6}.get(r[mu], 1e-6)
    a = agg.setdefault(r[kn][:90], [0, 0.0]); a[0] += 1; a[1] += v * scale
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f'{c:5d} {t:10.3f} ms  {k}')
PY
timeout 900 python bench.py > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/y_bench.err
