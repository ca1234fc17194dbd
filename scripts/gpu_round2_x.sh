#!/bin/bash
# round-2 session-2 batch 6: DRAM traffic of the banded Take / fused kernels at the benchmarked size, launch list of the bench
# command, full capture of the fused kernel (200M rows), reference arm
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none \
    -k regex:'take_kernel|take_cast_arith_kernel|take_validity_band_kernel' -c 8 --csv --log-file gpurun_out/take_traffic_r02b.csv \
    python scripts/take_traffic.py --variants valid,fused > gpurun_out/x_take_traffic.log 2>&1
tail -12 gpurun_out/take_traffic_r02b.csv | cut -c1-400
K='take_kernel|take_cast_arith_kernel|take_validity_band_kernel|map1_kernel|map2_kernel|bitmap_and_kernel|l2_demote_kernel|idx_locality_kernel'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 300 --csv \
    --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/x_bench_under_ncu.log 2>&1
tail -3 gpurun_out/x_bench_under_ncu.log | cut -c1-300
wc -l gpurun_out/launches_bench_r02.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:take_cast_arith_kernel -s 3 -c 1 \
    -o gpurun_out/fused_take_prof -f python bench.py --rows 200000000 --steps 1 --warmup 3 --no-configs > gpurun_out/fused_take_prof.log 2>&1
ncu -i gpurun_out/fused_take_prof.ncu-rep --page raw --csv > gpurun_out/fused_take_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/fused_take_prof.ncu-rep --page source --csv > gpurun_out/fused_take_prof_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/fused_take_prof_raw.csv
python scripts/ncu_source_top.py gpurun_out/fused_take_prof_source.csv 2>/dev/null | head -14
timeout 900 python bench.py --impl reference > gpurun_out/x_bench_ref.json 2> gpurun_out/x_bench_ref.err; echo "ref rc=$?"
tail -c 700 gpurun_out/x_bench_ref.json
