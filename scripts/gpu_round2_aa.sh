#!/bin/bash
# round-2 session-2 batch 9: suite (oversize guard test), f1 configs after the unmerged counting, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/aa_pytest.log
timeout 600 python bench_configs.py --only f1 > gpurun_out/aa_f1.jsonl 2> gpurun_out/aa_f1.err; echo "f1 rc=$?"
cat gpurun_out/aa_f1.jsonl | cut -c1-260
timeout 900 python bench.py > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/aa_bench.err
