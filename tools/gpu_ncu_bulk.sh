#!/bin/bash
# ncu --set full of the bulk kernels (propagate_universal, ephemeris, N-body) at their bench sizes
TAG=${1:-bulk}
mkdir -p gpurun_out
PERF_N=10000000 ncu --set full --clock-control none --import-source on -k regex:'propagate_universal_kernel' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_kepler -f python tools/gpu_perf_kepler.py > gpurun_out/${TAG}_ncu_kepler.log 2>&1; echo "ncu kepler rc=$?"
PERF_N=1000000 PERF_E=100 ncu --set full --clock-control none --import-source on -k regex:'ephemeris_twobody' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_eph -f python tools/gpu_perf_eph.py > gpurun_out/${TAG}_ncu_eph.log 2>&1; echo "ncu eph rc=$?"
PERF_N=500000 ncu --set full --clock-control none --import-source on -k regex:'propagate_nbody_kernel' --launch-skip 1 -c 1 \
    -o gpurun_out/${TAG}_nbody -f python tools/gpu_perf_nbody.py > gpurun_out/${TAG}_ncu_nbody.log 2>&1; echo "ncu nbody rc=$?"
tail -n 2 gpurun_out/${TAG}_ncu_kepler.log gpurun_out/${TAG}_ncu_eph.log gpurun_out/${TAG}_ncu_nbody.log
