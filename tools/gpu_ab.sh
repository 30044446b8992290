mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02x_pytest.log
python tools/gpu_perf_eph.py | tee gpurun_out/r02x_ab.log
python tools/gpu_perf_kepler.py | tee -a gpurun_out/r02x_ab.log
