mkdir -p gpurun_out
for v in "" outfit_b200/variants/lib_ung.so; do OUTFIT_B200_LIB=$v OUTFIT_B200_STREAMS=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB="; OUTFIT_B200_LIB=$v python tools/gpu_perf_eph.py; done | tee gpurun_out/r03a_ab.log
