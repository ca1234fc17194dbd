#!/bin/bash
# round 2, call N (the sweep below needs the createpolicy.range experiment that was measured and NOT kept --
# results in profiles/take_bitmap_prefix_pinning_r02.jsonl): bitmap prefix pinning sweep for Take; pinned result cache at the node level; take parity
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_pipeline.py tests/test_gpu_host_plugin.py "tests/test_gpu_parity.py::test_kat_take" -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -5 gpurun_out/n_pytest.log
: > gpurun_out/n_keep_sweep.jsonl
for mb in 0 24 40 64 96; do
  B2_TAKE_BITMAP_KEEP_MB=$mb timeout 300 python scripts/take_keep_sweep.py >> gpurun_out/n_keep_sweep.jsonl 2>> gpurun_out/n_keep_sweep.err
done
cat gpurun_out/n_keep_sweep.jsonl; tail -3 gpurun_out/n_keep_sweep.err
timeout 600 ./arrow_b200/lib/b200_host_test --bench-groupby 1000000000 10000000 4 > gpurun_out/n_node_bench.log 2>&1; echo "node bench rc=$?"
tail -3 gpurun_out/n_node_bench.log
