#!/bin/bash
# gpurun -- 'bash tools/profile_k2.sh <tag>': ncu capture of the policy-forward (tcgen05) and GAE kernels
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python tools/bench_ppo.py --envs 131072 --steps 8"
$CMD > gpurun_out/plain_k2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"policy_forward_kernel" -s 24 -c 6 -f -o gpurun_out/k2_$TAG $CMD > gpurun_out/ncu_k2_$TAG.log 2>&1
tail -1 gpurun_out/plain_k2_$TAG.log
