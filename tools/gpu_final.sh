#!/bin/bash
# Final evidence of a round (run on the GPU box via gpurun): full GPU test suite, smoke, default bench (both arms),
# launch list of one bench step, one `ncu --set full` capture per pairing kernel and of the calibration kernel.
# usage: tools/gpu_final.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -1 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
tools/gpu_ncu2.sh ${TAG} "k_pair_lines_duo k_miller k_fq_inv_batch k_fexp k_imad_peak" launches > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
