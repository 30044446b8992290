mkdir -p gpurun_out
TAG=${1:-ab}
for v in "" outfit_b200/variants/lib_sc6.so outfit_b200/variants/lib_sc8.so; do OUTFIT_B200_LIB=$v OUTFIT_B200_STREAMS=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB="; done | tee gpurun_out/${TAG}_ab.log
