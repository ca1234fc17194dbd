#!/bin/bash
# round 2, call O: cudaLimitMaxL2FetchGranularity sweep over the random-access kernels (experiment: no effect, knob and
# sweep script not kept; results in profiles/l2_fetch_granularity_sweep_r02.jsonl)
set -x
mkdir -p gpurun_out
: > gpurun_out/o_l2_sweep.jsonl
for g in unset 32 64 128; do
  if [ $g = unset ]; then timeout 400 python scripts/l2_granularity_sweep.py >> gpurun_out/o_l2_sweep.jsonl 2>> gpurun_out/o_l2_sweep.err
  else B2_L2_FETCH_GRANULARITY=$g timeout 400 python scripts/l2_granularity_sweep.py >> gpurun_out/o_l2_sweep.jsonl 2>> gpurun_out/o_l2_sweep.err; fi
done
cat gpurun_out/o_l2_sweep.jsonl; tail -5 gpurun_out/o_l2_sweep.err
