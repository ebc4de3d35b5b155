#!/bin/bash
# Run on the GPU box (via gpurun): kernel times against the batch size (whole waves of blocks or not), one sequence of
# kernels and the split form.   usage: tools/gpu_nscan.sh <tag> "<pairs list>"
TAG=$1; NS="$2"
mkdir -p gpurun_out
for sp in 1 2; do
for n in $NS; do
  BN_B200_SPLIT=$sp timeout 300 python bench.py --pairs $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n${n}_s${sp}.json 2> gpurun_out/${TAG}_n${n}_s${sp}.err
  python - gpurun_out/${TAG}_n${n}_s${sp}.json $n $sp <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']['kernels']
    n=int(sys.argv[2])
    print('n %6d split %s blocks %4d (%.2f waves) ms/step %.3f  per-pairing ns %.1f |'%(n,sys.argv[3],(n+19)//20,(n+19)//20/296,d['ms_per_step'],d['ms_per_step']*1e6/n),' '.join('%s %.3f'%(k,x['ms']) for k,x in r.items()), d.get('parity_failed',''))
except Exception as e:
    print('FAILED',sys.argv[1:],e)
PY
done; done
