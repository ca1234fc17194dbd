#!/bin/bash
# round 2, GPU call C2: ncu --set full of the compact partition passes (200M rows, cheap replay)
set -x
mkdir -p gpurun_out
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -o gpurun_out/$1 -f \
      python bench_configs.py --rows 200000000 --reps 1 --only c3 --fused-only > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$1_raw.csv > gpurun_out/$1_summary.txt 2>&1
  cat gpurun_out/$1_summary.txt
}
cap cpass1_prof 'compact_pass_kernel' 2
cap cpass0_prof 'compact_pass_kernel' 3
