#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_groupby_large.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden_fixtures.py -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -25 gpurun_out/g_pytest.log
timeout 600 python bench_configs.py --only c3 --reps 3 --fused-only > gpurun_out/g_c3.jsonl 2> gpurun_out/g_c3.err; cat gpurun_out/g_c3.jsonl; tail -c 300 gpurun_out/g_c3.err
timeout 900 ./arrow_b200/lib/b200_host_test --bench-groupby 1000000000 10000000 3 > gpurun_out/g_node_groupby.json 2> gpurun_out/g_node_groupby.err; echo "rc=$?"; cat gpurun_out/g_node_groupby.json; tail -c 400 gpurun_out/g_node_groupby.err
