#!/bin/bash
# round-2 session-2: per-kernel device times of the inner hash join (200M probe rows x 10M build keys) -- launch list only
mkdir -p gpurun_out
timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_join_r02.csv \
  -k regex:'probe_counts|emit_pairs|group_records|tile_scan|scan_tiles|tile_sums|grouper_|hashagg_|onesweep|sort_prepare|radix_|map1_kernel|filter_count|bitmap_and' \
  python bench_configs.py --rows 400000000 --reps 1 --only join > gpurun_out/ak_join_ncu.log 2>&1; echo "rc=$?"
wc -l gpurun_out/launches_join_r02.csv
