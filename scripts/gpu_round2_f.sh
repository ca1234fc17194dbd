#!/bin/bash
# round 2, GPU call F: dense group-by path -- parity tests, timing (1 vs 2 bands), launch list; node-level bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_groupby_large.py tests/test_gpu_binary.py tests/test_gpu_host_plugin.py -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -25 gpurun_out/f_pytest.log
timeout 600 python bench_configs.py --only c3 --reps 3 --fused-only > gpurun_out/f_c3.jsonl 2> gpurun_out/f_c3.err; cat gpurun_out/f_c3.jsonl; tail -c 500 gpurun_out/f_c3.err
B2_DENSE_BAND_MB=44 timeout 600 python bench_configs.py --only c3 --reps 3 --fused-only > gpurun_out/f_c3_2bands.jsonl 2>&1; cat gpurun_out/f_c3_2bands.jsonl
B2_GROUPBY_DENSE=0 timeout 600 python bench_configs.py --only c3 --reps 3 --fused-only > gpurun_out/f_c3_compact.jsonl 2>&1; cat gpurun_out/f_c3_compact.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'dense_|compact_|fused_' -c 60 --csv --log-file gpurun_out/launches_c3_dense_r02.csv python bench_configs.py --only c3 --reps 1 --fused-only > gpurun_out/f_ncu.log 2>&1
awk -F'","' '$(NF-2) ~ /gpu__time|dram__bytes_read|hit_rate/ {print $5, $(NF-2), $NF}' gpurun_out/launches_c3_dense_r02.csv | sed 's/b2:://g' | cut -c1-140 | tail -40
timeout 900 ./arrow_b200/lib/b200_host_test --bench-groupby 1000000000 10000000 3 > gpurun_out/f_node_groupby.json 2> gpurun_out/f_node_groupby.err; echo "rc=$?"; cat gpurun_out/f_node_groupby.json; tail -c 400 gpurun_out/f_node_groupby.err
