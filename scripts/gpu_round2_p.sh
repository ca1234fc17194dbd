#!/bin/bash
# round 2, call P: hash_sum consume through the dense path; c3 API path timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_groupby_large.py tests/test_gpu_parity.py tests/test_hash_aggregate_more.py tests/test_gpu_host_plugin.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -15 gpurun_out/p_pytest.log
timeout 600 python bench_configs.py --only c3 --reps 3 > gpurun_out/p_configs.jsonl 2> gpurun_out/p_configs.err; echo "configs rc=$?"
cat gpurun_out/p_configs.jsonl; tail -5 gpurun_out/p_configs.err
