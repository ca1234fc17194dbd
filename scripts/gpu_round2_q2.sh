#!/bin/bash
set -x
mkdir -p gpurun_out
./arrow_b200/lib/b200_host_test > gpurun_out/q_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/q_host_test.log; grep -v "^OK" gpurun_out/q_host_test.log | tail -15
