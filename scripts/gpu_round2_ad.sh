#!/bin/bash
# round-2 session-2 batch 11: host test (b200_hashjoin labels fix), join tests (validity pass), suite, bench
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/ad_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/ad_host_test.log; grep -v "^OK" gpurun_out/ad_host_test.log | tail -8 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/ad_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/ad_pytest.log
timeout 900 python bench_configs.py --only join > gpurun_out/ad_join.jsonl 2> gpurun_out/ad_join.err; echo "join rc=$?"
cat gpurun_out/ad_join.jsonl | cut -c1-300
timeout 900 python bench.py > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/ad_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
