mkdir -p gpurun_out
for st in ""; do echo "stages [$st]"; OUTFIT_B200_LIB=outfit_b200/variants/lib_dbg.so OUTFIT_B200_FG_STAGES="$st" python tools/gpu_stragglers.py 2>&1 | tail -5; done | tee gpurun_out/r01z2_sections.log
