#!/bin/bash
# Run on the GPU box (via gpurun): GPU parity tests, then a short bench per line-kernel mapping.
# usage: tools/gpu_ab3.sh <tag> "<labels>"   label = duo | solo | split<N> | <variant .so name> (libv_<name>.so, default mapping)
TAG=$1; LABELS="$2"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.txt
cat gpurun_out/${TAG}_pytest.txt
for v in $LABELS; do
  unset BN_B200_SO BN_B200_LINES BN_B200_SPLIT
  case $v in
    split*) export BN_B200_SPLIT=${v#split} ;;
    duo|solo) export BN_B200_LINES=$v ;;
    *) export BN_B200_SO=$PWD/bn_b200/libv_$v.so ;;
  esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "rc=$?" >> gpurun_out/${TAG}_bench_$v.err
  python - "$TAG" "$v" <<'PY'
import json,sys
tag,v=sys.argv[1:3]
try:
    d=json.loads(open('gpurun_out/%s_bench_%s.json'%(tag,v)).read().strip().splitlines()[-1])
    r=d.get('roofline',{})
    ks=r.get('kernels',{})
    print(v,'value %.0f e2e %.0f ms/step %.3f |'%(d['value'],d.get('e2e',{}).get('value',0),d['ms_per_step']),' '.join('%s %.3f'%(k,x['ms']) for k,x in ks.items()),'| frac %.3f whole %.3f pow %.0f g1mul %.0f fqmul %.3e'%(r.get('frac',0),r.get('whole_path_frac',0),r.get('fused_pairing_pow',{}).get('per_s',0),r.get('g1_scalar_mul',{}).get('per_s',0),r.get('fq_mul_chain',{}).get('fq_mul_per_s',0)), d.get('extras_error',''), d.get('parity_failed',''), d.get('parity',''))
except Exception as e:
    print(v,'FAILED',e); print(open('gpurun_out/%s_bench_%s.err'%(tag,v)).read()[-1500:])
PY
done
