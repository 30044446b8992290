#!/bin/bash
# ncu --set full of the three IOD kernels at the bench workload -> gpurun_out/<tag>_ncu_latest.json (bound to the kernel
# sources by sha) + the text summary; copy them to profiles/ncu_latest.json and profiles/<tag>_ncu_phases.txt
TAG=${1:-phases}
mkdir -p gpurun_out
python -c "import bench; print(bench.kernel_source_sha())" > gpurun_out/${TAG}_source_sha.txt
OUTFIT_B200_STREAMS=1 PERF_T=100000 PERF_PARITY=0 ncu --set full --clock-control none --import-source on -k regex:'roots_kernel|correct_kernel|score_kernel' \
    --launch-skip 3 -c 3 -o gpurun_out/${TAG}_phases -f python tools/gpu_perf.py > gpurun_out/${TAG}_ncu_phases.log 2>&1; echo "ncu phases rc=$?"
python tools/profile_report.py gpurun_out/${TAG}_phases.ncu-rep gpurun_out/${TAG}_ncu_phases.txt roots_kernel correct_kernel score_kernel > /dev/null 2>&1
python tools/ncu_latest.py gpurun_out/${TAG}_phases.ncu-rep gpurun_out/${TAG}_source_sha.txt c3_100k_x12 gpurun_out/${TAG}_ncu_latest.json
