#!/bin/bash
# round-2 session-2 batch 10: C++ host test (b200_hashjoin), join throughput, suite, bench
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/ac_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/ac_host_test.log; grep -v "^OK" gpurun_out/ac_host_test.log | tail -12 | cut -c1-400
timeout 900 python bench_configs.py --only join > gpurun_out/ac_join.jsonl 2> gpurun_out/ac_join.err; echo "join rc=$?"
cat gpurun_out/ac_join.jsonl | cut -c1-400; tail -3 gpurun_out/ac_join.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/ac_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/ac_pytest.log
timeout 900 python bench.py > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/ac_bench.err
