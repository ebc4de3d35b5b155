#!/bin/bash
# usage: tools/gpu_ncu.sh "<variant .so names or 'default'>"  -- one ncu --set full capture of k_miller_fexp per variant
mkdir -p gpurun_out
for v in $1; do
  if [ "$v" = "default" ]; then unset BN_B200_SO; else export BN_B200_SO=$PWD/bn_b200/$v; fi
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_miller_fexp -s 2 -c 1 -f -o gpurun_out/prof_miller_$v python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$v.log 2>&1
  tail -2 gpurun_out/ncu_$v.log
done
ls -la gpurun_out/*.ncu-rep
