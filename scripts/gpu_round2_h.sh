#!/bin/bash
# round 2: full GPU suite + the driver's bench command (N = 1) + reference arm; outputs for profiles/
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
timeout 1500 python bench.py > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "rc=$?"; tail -c 800 gpurun_out/h_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err; echo "rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; tail -2 gpurun_out/h_smoke.log
