#!/bin/bash
# gpurun -- 'bash tools/profile_k2_k5.sh <tag>': one ncu --set full capture each of the policy-forward kernel (K2, tcgen05) at
# 131 072 rows and of the update forward + backward kernel (K5, tcgen05) at 262 144 rows per minibatch, after the same command
# exited 0 without ncu; raw csv exported on the box (the reports stay in gpurun_out/ as well).
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python tools/bench_ppo.py --envs 131072 --steps 8"
$CMD > gpurun_out/plain_k2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"policy_forward_kernel" -s 24 -c 3 -f -o gpurun_out/k2_$TAG $CMD > gpurun_out/ncu_k2_$TAG.log 2>&1
$CMD > gpurun_out/plain_k5_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"update_fwdbwd_kernel" -s 4 -c 2 -f -o gpurun_out/k5_$TAG $CMD > gpurun_out/ncu_k5_$TAG.log 2>&1
tail -1 gpurun_out/plain_k2_$TAG.log | cut -c1-400
ls -la gpurun_out/k2_$TAG.ncu-rep gpurun_out/k5_$TAG.ncu-rep
