#!/bin/bash
# Run under gpurun on one B200: launch list + one full capture of the dominant kernel.
# Outputs land in gpurun_out/ (copy the summaries you want judged into profiles/).
set -x
K='take_kernel|map1_kernel|map2_kernel|bitmap_and_kernel'
# every launch of our kernels inside the default bench command, with device time
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the top kernel (smaller table so the replay save/restore stays cheap)
ncu --set full --clock-control none --import-source on -k regex:take_kernel -s 3 -c 1 \
    -o gpurun_out/take_prof -f python bench.py --rows 200000000 --steps 1 --warmup 3 > gpurun_out/take_prof.log 2>&1
ncu -i gpurun_out/take_prof.ncu-rep --page raw --csv > gpurun_out/take_prof_raw.csv 2>/dev/null
