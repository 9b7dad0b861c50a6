#!/bin/bash
# Run on the GPU box via gpurun: launch list + one full ncu capture of the env-step kernels the default configuration
# launches (qx::quadx_step_hot_kernel<REF, SHAPE, S1> + qx::quadx_reset_hot_kernel), each only after the same command
# exited 0 without ncu.   Usage: gpurun -- 'bash tools/profile_k1_r2.sh <tag> [envs]'
set -u
TAG=${1:-r2}
ENVS=${2:-1048576}
mkdir -p gpurun_out
CMD="python bench.py --steps 40 --warmup 3 --envs $ENVS --no-small --no-cpu-baseline --no-ppo --no-train"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'quadx_(step|reset)_hot' -s 40 -c 4 -f -o gpurun_out/k1_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/plain_$TAG.log
ls -la gpurun_out/k1_$TAG.ncu-rep
