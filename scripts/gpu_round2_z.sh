#!/bin/bash
# round-2 session-2 batch 8: wide-grouper tests (verification records), utf8 group-by timing, L2 demotion probe, suite, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grouper_wide.py tests/test_hash_aggregate_more.py -m gpu -x -q > gpurun_out/z_wide.log 2>&1; echo "wide rc=$?"
tail -4 gpurun_out/z_wide.log
timeout 900 python bench_configs.py --only c3u > gpurun_out/z_c3u.jsonl 2> gpurun_out/z_c3u.err; echo "c3u rc=$?"
cat gpurun_out/z_c3u.jsonl | cut -c1-300; tail -3 gpurun_out/z_c3u.err
for d in 0 1 0 1; do B2_L2_DEMOTE=$d timeout 600 python scripts/l2_demote_probe.py >> gpurun_out/z_l2_demote.jsonl 2>> gpurun_out/z_l2_demote.err; done
cat gpurun_out/z_l2_demote.jsonl; tail -3 gpurun_out/z_l2_demote.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/z_pytest.log
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/z_bench.err
