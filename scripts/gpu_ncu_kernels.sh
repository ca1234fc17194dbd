#!/bin/bash
# one `ncu --set full` capture per hot kernel at 200M rows (cheap replay), raw CSV pages for reading offline
set -x
cap() {  # name regex skip only
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -o gpurun_out/$1 -f \
      python bench_configs.py --rows 200000000 --reps 1 --only "$4" > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
}
cap filter_prof filter_compact_kernel 1 c1
cap onesweep_prof onesweep_kernel 9 c4
cap fused_prof fused_consume_kernel 1 c3
cap binfilter_prof 'filter_binary_kernel<long, 1' 1 c5
ls -la gpurun_out/*.ncu-rep
