#!/bin/bash
# one ncu --set full capture of each FitLSQ kernel (4-lane and one-lane) on the 100k x 12 batch
mkdir -p gpurun_out
OUTFIT_B200_LSQ_ONE_LANE=1 ncu --set full --clock-control none --import-source on -k regex:lsq_kernel -s 3 -c 1 -f -o gpurun_out/lsq_one python tools/gpu_perf_lsq.py > gpurun_out/lsq_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsq_quad -s 3 -c 1 -f -o gpurun_out/lsq_quad python tools/gpu_perf_lsq.py > gpurun_out/lsq_quad.log 2>&1
tail -3 gpurun_out/lsq_one.log gpurun_out/lsq_quad.log
