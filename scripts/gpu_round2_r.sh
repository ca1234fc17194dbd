#!/bin/bash
# round-2 session-2 batch 1: whole GPU suite on the restored tree + banded Take sweep + reduce profile + f1/cmp configs
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r_pytest.log
timeout 900 python scripts/take_band_sweep.py > gpurun_out/r_band_sweep.jsonl 2> gpurun_out/r_band_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r_band_sweep.jsonl; tail -3 gpurun_out/r_band_sweep.err
timeout 600 python bench_configs.py --only cmp,f1 > gpurun_out/r_configs.jsonl 2> gpurun_out/r_configs.err; echo "configs rc=$?"
cat gpurun_out/r_configs.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_kernel -c 1 -o gpurun_out/reduce_prof -f \
    python bench_configs.py --rows 200000000 --reps 1 --only cmp > gpurun_out/reduce_prof.log 2>&1
ncu -i gpurun_out/reduce_prof.ncu-rep --page raw --csv > gpurun_out/reduce_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/reduce_prof.ncu-rep --page source --csv > gpurun_out/reduce_prof_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/reduce_prof_raw.csv
