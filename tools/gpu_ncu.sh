#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> -- ncu launch list of one bench step + one --set full capture of each pairing kernel (default build)
mkdir -p gpurun_out
TAG=$1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
for k in k_pair_lines_duo k_miller k_fexp; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$k.log
done
ls -la gpurun_out/${TAG}_*
