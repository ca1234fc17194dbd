#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hash_join.py -x -q > gpurun_out/ab_join.log 2>&1; echo "join rc=$?"
tail -25 gpurun_out/ab_join.log | cut -c1-300
