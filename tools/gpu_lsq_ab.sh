#!/bin/bash
# A/B of the FitLSQ kernels: the 4-lane kernel at several residencies against the one-lane kernel; the sha1 of the
# result records + per-observation fit must be the same on every line.
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q -k "lsq" 2>&1 | tail -3
echo "== one lane =="; OUTFIT_B200_LSQ_ONE_LANE=1 python tools/gpu_perf_lsq.py
echo "== four lanes =="; python tools/gpu_perf_lsq.py


} > gpurun_out/lsq_ab.log 2>&1
tail -40 gpurun_out/lsq_ab.log
