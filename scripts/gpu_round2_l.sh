#!/bin/bash
# round 2, call L: direct-addressed Grouper -- parity, then c3 (Grouper path) and f1 timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grouper_direct.py tests/test_gpu_parity.py tests/test_hash_aggregate_more.py -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -15 gpurun_out/l_pytest.log
timeout 600 python bench_configs.py --only c3,f1 --reps 3 > gpurun_out/l_configs.jsonl 2> gpurun_out/l_configs.err; echo "configs rc=$?"
cat gpurun_out/l_configs.jsonl; tail -5 gpurun_out/l_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l_launches_grouper.csv python bench_configs.py --only c3 --reps 1 --rows 500000000 > gpurun_out/l_ncu.log 2>&1
grep -c . gpurun_out/l_launches_grouper.csv
