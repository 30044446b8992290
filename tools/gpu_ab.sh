mkdir -p gpurun_out
TAG=${1:-ab}
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
for st in "" "8,16"; do echo "stages [$st]"; OUTFIT_B200_LIB=outfit_b200/variants/lib_dbg.so OUTFIT_B200_FG_STAGES="$st" python tools/gpu_stragglers.py 2>&1 | tail -5; done | tee gpurun_out/${TAG}_stragglers.log
for st in "" "16" "8,16" "8,14,24" "7,10,14,20,30"; do
  echo "fg stages=[$st] single pass"; OUTFIT_B200_STREAMS=1 OUTFIT_B200_FG_STAGES="$st" PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB="
done | tee gpurun_out/${TAG}_ab.log
for st in "" "8,16" "8,14,24" "7,10,14,20,30"; do
  for ns in 2 8; do echo "fg stages=[$st] $ns passes"; OUTFIT_B200_STREAMS=$ns OUTFIT_B200_FG_STAGES="$st" PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "LIB="; done
done | tee -a gpurun_out/${TAG}_ab.log
