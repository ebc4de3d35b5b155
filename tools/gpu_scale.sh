#!/bin/bash
# usage: tools/gpu_scale.sh <tag> <N>   (inside `gpurun --gpus N`): bench.py at N ranks (torchrun) + the one-process multi-GPU C++ run
TAG=$1; N=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
python - "$TAG" "$N" <<'PY'
import json,sys
tag,n=sys.argv[1:3]
d=json.loads(open('gpurun_out/%s_bench_n%s.json'%(tag,n)).read().strip().splitlines()[-1])
print('N=%s value %.0f e2e %.0f ms/step %.3f gather=%s parity=%s'%(n,d['value'],d['e2e']['value'],d['ms_per_step'],d['gather'][:30],d.get('parity_checked')))
PY
libdir=$PWD/bn_b200
g++ -O1 -std=c++17 -o /tmp/api_main tests/cpp/api_main.cpp -L$libdir -lbn_b200 -Wl,-rpath,$libdir && /tmp/api_main --multi $N $((N * 16384)) /tmp/sample.bin | tee gpurun_out/${TAG}_cpp_multi_n$N.txt
