#!/bin/bash
# Round-2 profiling pass, run under gpurun on ONE GPU:   bash tools/profile_r02.sh [tag]
#  1. launch lists (ncu --metrics gpu__time_duration.sum) of `bench.py --config cN` -> gpurun_out/<tag>_launches_cN.csv
#  2. one `ncu --set full` capture of each config's dominant kernel(s)            -> gpurun_out/<tag>_full_cN.ncu-rep
# Numbers printed by runs under ncu are never bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --region-s 0.01 --no-e2e --no-cpu-baseline"
for c in c1 c2 c3 c4 c5; do
  L=""; [ $c = c5 ] && L="--lines 2097152"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/${TAG}_launches_$c.csv $B --config $c $L > gpurun_out/${TAG}_launches_$c.log 2>&1
done
declare -A K=( [c1]="split_kernel" [c2]="regex_tdfa" [c3]="split_kernel|ml_fused_kernel" [c4]="delim_kernel|regex_tdfa" [c5]="regex_tdfa_multi" )
for c in c1 c2 c3 c4 c5; do
  L=""; [ $c = c5 ] && L="--lines 1048576"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${K[$c]}" -s 4 -c 2 -f \
      -o gpurun_out/${TAG}_full_$c $B --config $c $L > gpurun_out/${TAG}_full_$c.log 2>&1
done
ls -la gpurun_out/${TAG}_*
