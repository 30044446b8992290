#!/bin/bash
# round 2, call 4: occupancy / tile A/B of the rewritten propagate and ephemeris kernels
mkdir -p gpurun_out
TAG=r2d
for v in default outfit_b200/variants/lib_prop_tpt3_bps5.so outfit_b200/variants/lib_prop_tpt6_bps5.so outfit_b200/variants/lib_prop_tpt8_bps5.so outfit_b200/variants/lib_prop_tpt4_bps6.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  echo "LIB=$v" | tee -a gpurun_out/${TAG}_kepler_ab.log
  python tools/gpu_perf_kepler.py 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_kepler_ab.log
done
for v in default outfit_b200/variants/lib_eph_bps4.so outfit_b200/variants/lib_eph_bps6.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  echo "LIB=$v" | tee -a gpurun_out/${TAG}_eph_ab.log
  python tools/gpu_perf_eph.py 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_eph_ab.log
done
unset OUTFIT_B200_LIB
python -m pytest tests -m gpu -q -k "propagate or ephemeris or restated or sweep" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
