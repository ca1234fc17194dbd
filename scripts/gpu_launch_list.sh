#!/bin/bash
# per-kernel device times of the non-headline configs (filter / group-by / sort / utf8) at full size
K='part_|preagg_|replay_|filter_|tile_scan|onesweep|radix_|sort_prepare|fused_|grouper_|hashagg_|take_|widen_|count_zero|bitmap_and'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 600 --csv \
    --log-file gpurun_out/launches_configs.csv python bench_configs.py --reps 1 --only ${1:-c1,c3,c4,c5} > gpurun_out/configs_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_configs.csv')) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0][:70]
    agg.setdefault(name, []).append(float(r[14]) / 1e6)
for k, v in agg.items():
    print(f"{k:72s} n={len(v):4d} total={sum(v):9.3f} ms  max={max(v):8.3f} ms")
PY
