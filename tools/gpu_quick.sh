#!/bin/bash
# Run on the GPU box (via gpurun): optional parity tests, then a short bench per library variant; one summary line each.
# usage: tools/gpu_quick.sh "<variant .so names or 'default'>" [test]
mkdir -p gpurun_out
if [ "$2" = "test" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
for v in $1; do
  if [ "$v" = "default" ]; then unset BN_B200_SO; else export BN_B200_SO=$PWD/bn_b200/$v; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2>gpurun_out/bench_$v.err; echo "rc=$?" >> gpurun_out/bench_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/bench_%s.json'%v).read().strip().splitlines()[-1])
    r=d['roofline']
    print('%-28s value %.0f e2e %.0f lines %.3f ms miller %.3f ms frac %.3f pow %.0f clocks %s'%(v,d['value'],d['e2e']['value'],r['kernel_ms']['k_pair_lines'],r['kernel_ms']['k_miller_fexp'],r['frac'],r['fused_pairing_pow']['per_s'],d['clocks']['sm_mhz']))
except Exception as e:
    print(v,'FAILED',e); print(open('gpurun_out/bench_%s.err'%v).read()[-1500:])
PY
done
