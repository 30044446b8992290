mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r03f_pytest.log
OUTFIT_B200_STREAMS=1 PERF_PARITY=1 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB=|parity" | tee gpurun_out/r03f_ab.log
python tools/gpu_perf_eph.py | tee -a gpurun_out/r03f_ab.log
