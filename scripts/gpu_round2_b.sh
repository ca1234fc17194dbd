#!/bin/bash
# round 2, GPU call B: compact group-by path -- parity tests, timing, launch list; multi-GPU legs at N=1
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_groupby_large.py tests/test_boolean_validity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -25 gpurun_out/b_pytest.log
timeout 600 python bench_configs.py --only c3 --reps 3 --fused-only > gpurun_out/b_c3.jsonl 2> gpurun_out/b_c3.err; echo "rc=$?"; cat gpurun_out/b_c3.jsonl; tail -c 800 gpurun_out/b_c3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3_r02.csv python bench_configs.py --only c3 --reps 1 --fused-only > gpurun_out/b_c3_ncu.log 2>&1
grep -E "compact|part_|preagg|fused_|emit|count_zero" gpurun_out/launches_c3_r02.csv | awk -F'","' '{print $5, $NF}' | tail -40
timeout 900 python bench.py --no-configs --steps 2 > gpurun_out/b_bench_multi.json 2> gpurun_out/b_bench_multi.err; echo "rc=$?"; tail -c 600 gpurun_out/b_bench_multi.err
