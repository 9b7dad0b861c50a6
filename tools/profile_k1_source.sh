#!/bin/bash
# GPU box: one ncu capture of the env-step kernel with per-instruction (SASS) stall samples, exported as csv on the box.
# Usage: gpurun -- 'QX_MERGED=1 QX_LANES=1 bash tools/profile_k1_source.sh <tag>'
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --set full --section SourceCounters --clock-control none -k regex:quadx_step_hot -s 6 -c 1 -f -o /tmp/k1_src_$TAG python tools/perstep.py 1048576 10 > gpurun_out/ncu_src_$TAG.log 2>&1
ncu -i /tmp/k1_src_$TAG.ncu-rep --page raw --csv > gpurun_out/k1_src_$TAG.raw.csv 2>/dev/null
ncu -i /tmp/k1_src_$TAG.ncu-rep --page source --print-source sass --csv > gpurun_out/k1_src_$TAG.sass.csv 2>/dev/null
ls -la gpurun_out/k1_src_$TAG.*; tail -2 gpurun_out/ncu_src_$TAG.log | cut -c1-200
