#!/bin/bash
# Run on the GPU box via gpurun: launch list + one full ncu capture of the env-step kernel.
# Usage: gpurun -- 'bash tools/profile_k1.sh <tag> [envs]'
set -u
TAG=${1:-r1}
ENVS=${2:-1048576}
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --envs $ENVS --no-small --no-cpu-baseline --no-ppo --no-train"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:quadx_step_kernel -s 5 -c 4 -f -o gpurun_out/k1_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
ls -la gpurun_out/
