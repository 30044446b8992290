#!/bin/bash
# A/B of differently built copies of the library on the GPU box (run under gpurun).  Build the variants
# in-tree first, e.g.  make -C outfit_b200/csrc OUT=../variants/lib_x.so EXTRA=-DSOME_SWITCH=1
# usage: tools/gpu_ab.sh outfit_b200/variants/lib_a.so outfit_b200/variants/lib_b.so ...
# Each variant runs the IOD perf script with the oracle parity check on a small batch (PERF_T, default 5000).
for v in "$@"; do
  OUTFIT_B200_LIB=$v PERF_T=${PERF_T:-5000} OUTFIT_B200_STREAMS=1 PERF_PARITY=1 python tools/gpu_perf.py 2>&1 | grep -E "LIB=|parity|phases"
done
