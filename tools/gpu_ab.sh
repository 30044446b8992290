mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r01e_pytest_gpu.log 2>&1; tail -5 gpurun_out/r01e_pytest_gpu.log
for v in "" outfit_b200/variants/lib_bps3.so outfit_b200/variants/lib_bps5.so outfit_b200/variants/lib_bps6.so; do
  OUTFIT_B200_LIB=$v PERF_PARITY=$([ -z "$v" ] && echo 1 || echo 0) python tools/gpu_perf.py 2>&1 | tail -4
done | tee gpurun_out/r01e_ab.log
PERF_COUNT=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | tail -3 | tee -a gpurun_out/r01e_ab.log
