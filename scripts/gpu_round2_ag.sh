#!/bin/bash
# round-2 session-2 batch 14: join with 32-bit group records (tests, host test, throughput), whole suite, bench line, smoke
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/ag_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/ag_host_test.log; grep -v "^OK" gpurun_out/ag_host_test.log | tail -6 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ag_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/ag_pytest.log | cut -c1-300
timeout 900 python bench_configs.py --only join > gpurun_out/ag_join.jsonl 2> gpurun_out/ag_join.err; echo "join rc=$?"
cat gpurun_out/ag_join.jsonl | cut -c1-300
timeout 900 python bench.py > gpurun_out/ag_bench.json 2> gpurun_out/ag_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/ag_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
