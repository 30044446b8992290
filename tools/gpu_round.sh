#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of the three numeric kernels.  Outputs land in gpurun_out/<tag>_*.
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${TAG}_bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
[ "${SKIP_NCU_IOD:-0}" = "1" ] || OUTFIT_B200_STREAMS=1 PERF_T=30000 PERF_PARITY=0 ncu --set full --clock-control none --import-source on -k regex:'roots_kernel|correct_kernel|score_kernel' \
    --launch-skip 3 -c 3 -o gpurun_out/${TAG}_phases -f python tools/gpu_perf.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:'lsq_quad_kernel' --launch-skip 1 -c 1 \
    -o gpurun_out/${TAG}_lsq -f python tools/gpu_perf_lsq.py 30000 > gpurun_out/${TAG}_ncu_lsq.log 2>&1; echo "ncu lsq rc=$?"
fi
