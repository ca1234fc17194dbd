#!/bin/bash
# round 2, GPU call A: full GPU test suite, the new bench line (small then full size), Take traffic split at 1B rows
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g | head -2 >> gpurun_out/a_gpu.txt; lscpu | grep -E "Model name|NUMA" >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --rows 50000000 --steps 2 > gpurun_out/a_bench_small.json 2> gpurun_out/a_bench_small.err; echo "rc=$?"
tail -c 1500 gpurun_out/a_bench_small.err
timeout 1200 python bench.py > gpurun_out/a_bench_full.json 2> gpurun_out/a_bench_full.err; echo "rc=$?"
tail -c 1500 gpurun_out/a_bench_full.err
timeout 600 python bench.py --impl reference > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
    --clock-control none -k regex:take_kernel -c 2 --csv --log-file gpurun_out/take_traffic_r02.csv python scripts/take_traffic.py > gpurun_out/a_take_traffic.log 2>&1
tail -3 gpurun_out/take_traffic_r02.csv
