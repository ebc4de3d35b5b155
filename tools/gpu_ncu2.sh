#!/bin/bash
# usage: tools/gpu_ncu2.sh <tag> "<kernel names>" [launches]  -- one `ncu --set full` capture per named pairing kernel
# inside `python bench.py --steps 1 --warmup 3` (2^14 pairings), optionally the launch list of one bench step first.
# The captures run with BN_B200_SPLIT=1: the FULL-SIZE kernels (one sequence per call), as in bench.py's per-kernel
# timing pass that the roofline numbers come from; the launch list shows the calls as the product runs them (two sub-batches).
mkdir -p gpurun_out
TAG=$1; KERNELS="$2"
if [ "$3" = "launches" ]; then
  BN_B200_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
fi
for k in $KERNELS; do
  BN_B200_SPLIT=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$k.log
done
ls -la gpurun_out/${TAG}_*
