#!/bin/bash
# one ncu --set full capture of the FitLSQ kernel on the 100k x 12 batch
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lsq_quad -s 3 -c 1 -f -o gpurun_out/lsq_quad python tools/gpu_perf_lsq.py > gpurun_out/lsq_quad.log 2>&1
tail -n 3 gpurun_out/lsq_quad.log
