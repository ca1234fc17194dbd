#!/bin/bash
# round-2 session-2 batch 5: suite after the reduce rewrite / select_k fix, reduce timing + profile, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/v_pytest.log
timeout 600 python bench_configs.py --only cmp,f1 > gpurun_out/v_configs.jsonl 2> gpurun_out/v_configs.err; echo "configs rc=$?"
cat gpurun_out/v_configs.jsonl
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'b2::reduce_kernel' -c 1 -o gpurun_out/reduce_prof_b -f \
    python bench_configs.py --rows 200000000 --reps 1 --only cmp > gpurun_out/reduce_prof_b.log 2>&1
ncu -i gpurun_out/reduce_prof_b.ncu-rep --page raw --csv > gpurun_out/reduce_prof_b_raw.csv 2>/dev/null
ncu -i gpurun_out/reduce_prof_b.ncu-rep --page source --csv > gpurun_out/reduce_prof_b_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/reduce_prof_b_raw.csv
python scripts/ncu_source_top.py gpurun_out/reduce_prof_b_source.csv 2>/dev/null | head -16
timeout 900 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/v_bench.err
