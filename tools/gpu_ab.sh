#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, then bench for each library variant, then ncu captures.
# usage: tools/gpu_ab.sh "<variant .so names or 'default'>" [ncu]
mkdir -p gpurun_out
VARIANTS="$1"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for spec in $VARIANTS; do
  v=${spec%%:*}; lines=${spec##*:}; [ "$lines" = "$spec" ] && lines=duo
  export BN_B200_LINES=$lines
  if [ "$v" = "default" ]; then unset BN_B200_SO; else export BN_B200_SO=$PWD/bn_b200/$v; fi
  extra="--no-cpu-baseline"; [ "$spec" = "default" ] && extra=""
  v=${v}_$lines
  timeout 600 python bench.py --steps 8 --warmup 3 $extra > gpurun_out/bench_$v.log 2>gpurun_out/bench_$v.err; echo "rc=$?" >> gpurun_out/bench_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/bench_%s.log'%v).read().strip().splitlines()[-1])
    r=d['roofline']
    print(v,'value %.0f e2e %.0f ms/step %.3f lines %.3f ms miller %.3f ms fexp %.3f ms imad_peak %.2f T frac %.3f whole %.3f fqmul %.3e (%.2f) clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['kernel_ms']['k_pair_lines'],r['kernel_ms']['k_miller'],r['kernel_ms']['k_fexp'],r['peak'],r['frac'],r['whole_path_frac'],r['fq_mul_chain']['fq_mul_per_s'],r['fq_mul_chain']['imad_frac'],d['clocks']))
except Exception as e:
    print(v,'FAILED',e); print(open('gpurun_out/bench_%s.err'%v).read()[-1500:])
PY
done
unset BN_B200_SO BN_B200_LINES
if [ "$2" = "ncu" ]; then
  tools/gpu_ncu.sh ab   # launch list of one bench step + one --set full capture of k_pair_lines_duo, k_miller, k_fexp
fi
