#!/bin/bash
# round 2, call 3: sanitizer on the rewritten kernels, all GPU tests + sweep, A/B of the propagate / ephemeris rewrites
mkdir -p gpurun_out
TAG=r2c
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/${TAG}_racecheck.log
for v in outfit_b200/variants/lib_r2b.so default outfit_b200/variants/lib_prop_tpt2_bps4.so outfit_b200/variants/lib_prop_tpt8_bps4.so outfit_b200/variants/lib_prop_tpt8_bps3.so outfit_b200/variants/lib_prop_tpt4_bps3.so outfit_b200/variants/lib_prop_tpt4_bps5.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  echo "LIB=$v" | tee -a gpurun_out/${TAG}_kepler_ab.log
  python tools/gpu_perf_kepler.py 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_kepler_ab.log
done
for v in outfit_b200/variants/lib_r2b.so default; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  echo "LIB=$v" | tee -a gpurun_out/${TAG}_eph_ab.log
  python tools/gpu_perf_eph.py 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_eph_ab.log
done
unset OUTFIT_B200_LIB
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
PERF_N=10000000 ncu --set full --clock-control none --import-source on -k regex:'propagate_universal_kernel' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_kepler -f python tools/gpu_perf_kepler.py > gpurun_out/${TAG}_ncu_kepler.log 2>&1; echo "ncu kepler rc=$?"
PERF_N=1000000 PERF_E=100 ncu --set full --clock-control none --import-source on -k regex:'ephemeris_twobody' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_eph -f python tools/gpu_perf_eph.py > gpurun_out/${TAG}_ncu_eph.log 2>&1; echo "ncu eph rc=$?"
