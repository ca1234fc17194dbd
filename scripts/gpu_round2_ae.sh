#!/bin/bash
# round-2 session-2 batch 12: host test (join schema suffixes), whole suite, bench, smoke
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/ae_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/ae_host_test.log; grep -v "^OK" gpurun_out/ae_host_test.log | tail -8 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ae_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/ae_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/ae_bench.json 2> gpurun_out/ae_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/ae_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
