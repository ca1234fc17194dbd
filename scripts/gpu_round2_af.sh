#!/bin/bash
# round-2 session-2 batch 13: FULL OUTER join (C-ABI tests + Acero node), whole suite
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/af_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/af_host_test.log; grep -v "^OK" gpurun_out/af_host_test.log | tail -8 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/af_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/af_pytest.log | cut -c1-300
