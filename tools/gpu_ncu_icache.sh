#!/bin/bash
# usage: tools/gpu_ncu_icache.sh "<variants>" <kernel-regex>  -- instruction-cache (GCC = L1.5) counters + stall reasons of one kernel
mkdir -p gpurun_out
M=gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,gcc__gcc2xbar_requests_type_instruction.sum,gcc__average_cache_request_type_instruction_hit_rate.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,idc__requests.sum,idc__requests_lookup_miss.sum
for v in $1; do
  if [ "$v" = "default" ]; then unset BN_B200_SO; else export BN_B200_SO=$PWD/bn_b200/$v; fi
  timeout 300 ncu --metrics $M --clock-control none -k regex:$2 -s 2 -c 1 --csv --log-file gpurun_out/icache_$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  echo "== $v"; python - "$v" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/icache_%s.csv'%sys.argv[1])) if len(r)>10]
for r in rows[1:]: print("%-80s %s"%(r[-3][:80], r[-1]))
PY
done
