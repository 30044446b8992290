mkdir -p gpurun_out
for pr in 1 0; do echo "score pruning=$pr"; PERF_T=20000 PERF_COUNT=1 OUTFIT_B200_SCORE_PRUNING=$pr OUTFIT_B200_STREAMS=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB=|counters"; done | tee gpurun_out/r02n_ab.log
