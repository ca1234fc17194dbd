#!/bin/bash
# round-2 session-2 batch 3: whole suite + host test (new Acero nodes, count_distinct), bench line, configs, reduce profile
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/t_pytest.log
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/t_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/t_host_test.log; grep -v "^OK" gpurun_out/t_host_test.log | tail -15
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/t_bench.json; tail -5 gpurun_out/t_bench.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'b2::reduce_kernel' -c 1 -o gpurun_out/reduce_prof -f \
    python bench_configs.py --rows 200000000 --reps 1 --only cmp > gpurun_out/reduce_prof.log 2>&1
ncu -i gpurun_out/reduce_prof.ncu-rep --page raw --csv > gpurun_out/reduce_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/reduce_prof.ncu-rep --page source --csv > gpurun_out/reduce_prof_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/reduce_prof_raw.csv
python scripts/ncu_source_top.py gpurun_out/reduce_prof_source.csv 2>/dev/null | head -30
tail -5 gpurun_out/reduce_prof.log
