#!/bin/bash
# round 2, call 1: GPU tests (incl. the full parity sweep), bench at HEAD, correct_kernel occupancy / block-size A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest_gpu.log
tail -15 gpurun_out/r2a_pytest_gpu.log
for v in default outfit_b200/variants/lib_bps5.so outfit_b200/variants/lib_ct64_bps8.so outfit_b200/variants/lib_ct64_bps10.so outfit_b200/variants/lib_ct32_bps16.so outfit_b200/variants/lib_score5.so outfit_b200/variants/lib_score4.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  PERF_T=100000 OUTFIT_B200_STREAMS=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "LIB=|phases" | tee -a gpurun_out/r2a_correct_ab.log
  PERF_T=100000 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "LIB=" | tee -a gpurun_out/r2a_correct_ab.log
done
unset OUTFIT_B200_LIB
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r2a_bench.json
