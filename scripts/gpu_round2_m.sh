#!/bin/bash
# round 2, call M: fused mean, grouper minmax/count tweaks; node-level group-by bench; f1
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grouper_direct.py tests/test_gpu_parity.py tests/test_gpu_host_plugin.py -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -15 gpurun_out/m_pytest.log
timeout 600 python bench_configs.py --only c3,f1 --reps 3 > gpurun_out/m_configs.jsonl 2> gpurun_out/m_configs.err; echo "configs rc=$?"
cat gpurun_out/m_configs.jsonl; tail -5 gpurun_out/m_configs.err
timeout 600 ./arrow_b200/lib/b200_host_test --bench-groupby 1000000000 10000000 4 > gpurun_out/m_node_bench.log 2>&1; echo "node bench rc=$?"
tail -8 gpurun_out/m_node_bench.log
