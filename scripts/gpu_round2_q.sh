#!/bin/bash
# round 2, call Q: if_else, boolean/validity in the C++ registry, multi-key sort -- parity on the GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_if_else.py tests/test_sort_multi_key.py tests/test_gpu_host_plugin.py tests/test_boolean_validity.py tests/test_cabi.py -m gpu -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -15 gpurun_out/q_pytest.log
./arrow_b200/lib/b200_host_test > gpurun_out/q_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/q_host_test.log; grep -v "^OK" gpurun_out/q_host_test.log | tail -15
