#!/bin/bash
# round 2, GPU call I: new tests, fused-kernel DRAM traffic at 1B rows, profiles of the kernels as committed, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fused_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_groupby_large.py -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -12 gpurun_out/i_pytest.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:'take_kernel|take_cast_arith' -c 3 --csv --log-file gpurun_out/take_traffic_r02.csv python scripts/take_traffic.py > gpurun_out/i_take_traffic.log 2>&1
grep -E "take" gpurun_out/take_traffic_r02.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-150
cap() {  # name regex skip only
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -o gpurun_out/$1 -f \
      python bench_configs.py --rows 200000000 --reps 1 --only "$4" --fused-only > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$1_raw.csv > gpurun_out/$1_summary.txt 2>&1
  cat gpurun_out/$1_summary.txt
}
cap onesweep_prof_r02 onesweep_kernel 9 c4
cap dense_prof_r02 dense_consume_kernel 2 c3
timeout 1500 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/i_bench.err
