#!/bin/bash
# GPU box: time every env-step kernel variant, then one ncu --set full capture of three of them.
# Usage: gpurun -- 'bash tools/profile_k1_variants.sh <tag>'
TAG=${1:-r2}
mkdir -p gpurun_out
python tools/k1_variants.py 1048576 200 > gpurun_out/k1_variants_$TAG.txt 2>&1
cat gpurun_out/k1_variants_$TAG.txt | cut -c1-200
for V in one_env pair_packed_shape0 pair_scalar_shape0; do
  K1_ONLY=$V ncu --set full --clock-control none --import-source on -k regex:quadx_step -s 8 -c 2 -f -o gpurun_out/k1_${V}_$TAG python tools/k1_variants.py 1048576 6 > gpurun_out/ncu_${V}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${V}_$TAG.log | cut -c1-200
done
