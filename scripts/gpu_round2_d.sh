#!/bin/bash
# round 2, GPU call D: full GPU suite, compact group-by + staged utf8 copy timings, node-level group-by bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -15 gpurun_out/d_pytest.log
timeout 600 python bench_configs.py --only c3,c5 --reps 3 --fused-only > gpurun_out/d_c3.jsonl 2> gpurun_out/d_c3.err; cat gpurun_out/d_c3.jsonl
timeout 600 python bench_configs.py --only c5 --reps 3 > gpurun_out/d_c5.jsonl 2> gpurun_out/d_c5.err; echo "rc=$?"; cat gpurun_out/d_c5.jsonl; tail -c 600 gpurun_out/d_c5.err
B2_BINARY_STAGED=0 timeout 600 python bench_configs.py --only c5 --reps 3 > gpurun_out/d_c5_chunk.jsonl 2>&1; cat gpurun_out/d_c5_chunk.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'compact|filter_binary|tile_scan|filter_count|fused_' -c 60 --csv --log-file gpurun_out/launches_c3c5_r02.csv python bench_configs.py --only c3,c5 --reps 1 --fused-only > gpurun_out/d_ncu.log 2>&1
awk -F'","' '{print $5, $NF}' gpurun_out/launches_c3c5_r02.csv | sed 's/b2:://g' | cut -c1-110 | tail -24
timeout 900 ./arrow_b200/lib/b200_host_test --bench-groupby 1000000000 10000000 3 > gpurun_out/d_node_groupby.json 2> gpurun_out/d_node_groupby.err; echo "rc=$?"; cat gpurun_out/d_node_groupby.json; tail -c 400 gpurun_out/d_node_groupby.err
