mkdir -p gpurun_out
for v in outfit_b200/variants/lib_base.so "" outfit_b200/variants/lib_bps5.so outfit_b200/variants/lib_bps6.so; do OUTFIT_B200_LIB=$v OUTFIT_B200_STREAMS=1 PERF_PARITY=1 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB=|parity"; done | tee gpurun_out/r02r_ab.log
