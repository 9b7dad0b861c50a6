#!/bin/bash
# GPU box: per-step anatomy of the merged / static / two-launch env step, then one ncu --set full capture of each,
# exported as raw csv on the box (the reports themselves are too large to bring back three at a time).
# Usage: gpurun -- 'bash tools/profile_k1_merged.sh <tag> [modes]'
TAG=${1:-r2}
MODES=${2:-"0 1 2"}
mkdir -p gpurun_out
for M in $MODES; do
  QX_MERGED=$M python tools/perstep.py 1048576 24 2>&1 | cut -c1-900 > gpurun_out/perstep_${TAG}_m$M.txt
  QX_MERGED=$M ncu --set full --clock-control none -k regex:quadx_step_hot -s 6 -c 1 -f -o /tmp/k1_m${M}_$TAG python tools/perstep.py 1048576 10 > gpurun_out/ncu_m${M}_$TAG.log 2>&1
  ncu -i /tmp/k1_m${M}_$TAG.ncu-rep --page raw --csv > gpurun_out/k1_m${M}_$TAG.raw.csv 2>/dev/null
  tail -2 gpurun_out/ncu_m${M}_$TAG.log | cut -c1-200
done
