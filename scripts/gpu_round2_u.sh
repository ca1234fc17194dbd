#!/bin/bash
# round-2 session-2 batch 4: host test (Acero map nodes, select_k, count_distinct), suite, take demote / locality sweep
set -x
mkdir -p gpurun_out
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/u_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/u_host_test.log; grep -v "^OK" gpurun_out/u_host_test.log | tail -15
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/u_pytest.log
timeout 900 python scripts/take_band_sweep.py --bands 0nd,0,64nd,64 > gpurun_out/u_band_sweep.jsonl 2> gpurun_out/u_band_sweep.err; echo "sweep rc=$?"
cat gpurun_out/u_band_sweep.jsonl; tail -3 gpurun_out/u_band_sweep.err
