#!/bin/bash
# round 2, call 5: higher-occupancy variants (propagate, ephemeris, scorer)
mkdir -p gpurun_out
TAG=r2e
for v in default outfit_b200/variants/lib_prop_tpt6_bps6.so outfit_b200/variants/lib_prop_tpt4_bps7.so outfit_b200/variants/lib_prop_tpt4_bps8.so outfit_b200/variants/lib_prop_tpt5_bps7.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  echo "LIB=$v" | tee -a gpurun_out/${TAG}_kepler_ab.log
  python tools/gpu_perf_kepler.py 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_kepler_ab.log
done
for v in default outfit_b200/variants/lib_eph_bps7.so outfit_b200/variants/lib_eph_bps8.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  echo "LIB=$v" | tee -a gpurun_out/${TAG}_eph_ab.log
  python tools/gpu_perf_eph.py 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_eph_ab.log
done
for v in default outfit_b200/variants/lib_score7.so outfit_b200/variants/lib_score8.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  PERF_T=100000 OUTFIT_B200_STREAMS=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "LIB=|phases" | tee -a gpurun_out/${TAG}_score_ab.log
done
