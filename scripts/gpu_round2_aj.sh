#!/bin/bash
# round-2 session-2 batch 17: the whole GPU suite on the final tree
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x > gpurun_out/aj_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/aj_pytest.log | cut -c1-200
