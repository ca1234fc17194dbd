#!/bin/bash
# round-2 session-2 batch 15 (last): host test incl. the utf8-key aggregate plan, whole suite, bench line, smoke
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/ah_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/ah_host_test.log; grep -v "^OK" gpurun_out/ah_host_test.log | tail -6 | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ah_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/ah_pytest.log | cut -c1-300
timeout 900 python bench.py --no-configs > gpurun_out/ah_bench.json 2> gpurun_out/ah_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/ah_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
