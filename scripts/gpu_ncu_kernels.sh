#!/bin/bash
# `ncu --set full` captures of hot kernels at 200M rows (cheap replay); raw + source CSV pages for reading offline.
# usage: gpu_ncu_kernels.sh [name regex skip only]...   (no args = the default set)
set -x
cap() {  # name regex skip only
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -o gpurun_out/$1 -f \
      python bench_configs.py --rows 200000000 --reps 1 --only "$4" > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
}
if [ $# -ge 4 ]; then
  while [ $# -ge 4 ]; do cap "$1" "$2" "$3" "$4"; shift 4; done
else
  cap filter_prof filter_compact_kernel 1 c1
  cap onesweep_prof onesweep_kernel 9 c4
  cap partpass_prof 'part_pass_kernel<0' 1 c3
  cap preagg_prof preagg_kernel 1 c3
  cap binfilter_prof 'filter_binary_kernel<long, 1' 1 c5
fi
ls -la gpurun_out/*.ncu-rep
