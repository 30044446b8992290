mkdir -p gpurun_out
TAG=${1:-ab}
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
for ns in 1 2 4 8; do echo "$ns passes"; OUTFIT_B200_STREAMS=$ns PERF_E2E=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "phases|LIB=|e2e host"; done | tee gpurun_out/${TAG}_ab.log
