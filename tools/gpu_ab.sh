mkdir -p gpurun_out
TAG=${1:-ab}
for v in "" outfit_b200/variants/lib_pred.so; do OUTFIT_B200_LIB=$v OUTFIT_B200_STREAMS=1 PERF_PARITY=1 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB=|parity"; OUTFIT_B200_LIB=$v python tools/gpu_perf_kepler.py; done | tee gpurun_out/${TAG}_ab.log
