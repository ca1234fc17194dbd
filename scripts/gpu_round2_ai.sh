#!/bin/bash
# round-2 session-2 batch 16: C++ host test after the hash_count_distinct output-type fix
set -x
mkdir -p gpurun_out
timeout 300 ./arrow_b200/lib/b200_host_test > gpurun_out/ai_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/ai_host_test.log; grep -v "^OK" gpurun_out/ai_host_test.log | tail -6 | cut -c1-600
