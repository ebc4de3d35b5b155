#!/bin/bash
# Final evidence of a round (run on the GPU box via gpurun; two calls, gpurun_out/ is limited to 64 MiB per call):
#   tools/gpu_final.sh <tag> run   full GPU test suite, smoke, default bench (both arms), launch list of the bench as shipped
#   tools/gpu_final.sh <tag> ncu   launch list with one sequence of kernels per call, one `ncu --set full` capture per
#                                  pairing kernel and of the calibration kernel
TAG=$1
mkdir -p gpurun_out
if [ "$2" = "run" ]; then
  timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_pytest_gpu.txt
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -1 gpurun_out/${TAG}_smoke.txt
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
  # the calls as the product runs them (2^14 pairings: two sub-batches; ncu serialises kernels that overlap in a real run)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_split.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
  ls -la gpurun_out/
else
  # launch list of one sequence of full-size kernels per call (BN_B200_SPLIT=1, comparable with bench.py's per-kernel events), then the captures
  tools/gpu_ncu2.sh ${TAG} "k_pair_lines_duo k_miller k_fq_inv_batch k_fexp k_imad_peak" launches > gpurun_out/${TAG}_ncu.log 2>&1
  # digest the captures here (the .ncu-rep files together exceed the 64 MiB that travel back): per-kernel summary, raw page,
  # instruction mix -> ncu_kernels.json; keep only the two big kernels' reports
  R=gpurun_out/${TAG}_prof
  python tools/ncu_kernels.py ${TAG} gpurun_out/${TAG}_ncu_kernels.json k_pair_lines_duo=${R}_k_pair_lines_duo.ncu-rep k_miller=${R}_k_miller.ncu-rep \
      k_fq_inv_batch=${R}_k_fq_inv_batch.ncu-rep k_fexp=${R}_k_fexp.ncu-rep k_imad_peak=${R}_k_imad_peak.ncu-rep > gpurun_out/${TAG}_ncu_kernels.txt 2>&1
  python tools/ncu_summary.py ${R}_k_pair_lines_duo.ncu-rep ${R}_k_miller.ncu-rep ${R}_k_fq_inv_batch.ncu-rep ${R}_k_fexp.ncu-rep ${R}_k_imad_peak.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
  for k in k_pair_lines_duo k_miller k_fq_inv_batch k_fexp k_imad_peak; do ncu -i ${R}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_$k.csv 2>/dev/null; done
  rm -f ${R}_k_pair_lines_duo.ncu-rep ${R}_k_fq_inv_batch.ncu-rep ${R}_k_imad_peak.ncu-rep
  cat gpurun_out/${TAG}_ncu_kernels.txt; du -sh gpurun_out
fi
