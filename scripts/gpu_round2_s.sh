#!/bin/bash
# round-2 session-2 batch 2: wide / utf8 grouper tests, C++ host test, Take band K=2, reduce profile (own kernel this time)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grouper_wide.py -x -q > gpurun_out/s_wide.log 2>&1; echo "wide rc=$?"
tail -25 gpurun_out/s_wide.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_grouper_wide.py > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s_pytest.log
timeout 600 ./arrow_b200/lib/b200_host_test > gpurun_out/s_host_test.log 2>&1; echo "host test rc=$?"
grep -c "^OK" gpurun_out/s_host_test.log; grep -v "^OK" gpurun_out/s_host_test.log | tail -15
timeout 900 python scripts/take_band_sweep.py --bands 0,48,64,80 > gpurun_out/s_band_sweep.jsonl 2> gpurun_out/s_band_sweep.err; echo "sweep rc=$?"
cat gpurun_out/s_band_sweep.jsonl; tail -3 gpurun_out/s_band_sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'reduce_kernel<long' -c 1 -o gpurun_out/reduce_prof -f \
    python bench_configs.py --rows 200000000 --reps 1 --only cmp > gpurun_out/reduce_prof.log 2>&1
ncu -i gpurun_out/reduce_prof.ncu-rep --page raw --csv > gpurun_out/reduce_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/reduce_prof.ncu-rep --page source --csv > gpurun_out/reduce_prof_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/reduce_prof_raw.csv
python scripts/ncu_source_top.py gpurun_out/reduce_prof_source.csv 2>/dev/null | head -30
