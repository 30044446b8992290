#!/bin/bash
# One gpurun call for the round's evidence at HEAD: GPU tests (incl. the full parity sweep), the bench line + the reference
# arm, `ncu --set full` of the three IOD kernels at the bench workload (-> profiles/ncu_latest.json, bound to the kernel
# sources by their sha), of the bulk kernels, and the ncu launch list of the bench command.
TAG=${1:-r2g}
mkdir -p gpurun_out
python -c "import bench; print(bench.kernel_source_sha())" > gpurun_out/${TAG}_source_sha.txt
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench.err; echo "reference arm rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
if [ "${SKIP_PHASES:-0}" != "1" ]; then
OUTFIT_B200_STREAMS=1 PERF_T=100000 PERF_PARITY=0 ncu --set full --clock-control none --import-source on -k regex:'roots_kernel|correct_kernel|score_kernel' \
    --launch-skip 3 -c 3 -o gpurun_out/${TAG}_phases -f python tools/gpu_perf.py > gpurun_out/${TAG}_ncu_phases.log 2>&1; echo "ncu phases rc=$?"
fi
PERF_N=10000000 ncu --set full --clock-control none --import-source on -k regex:'propagate_universal_kernel' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_kepler -f python tools/gpu_perf_kepler.py > gpurun_out/${TAG}_ncu_kepler.log 2>&1; echo "ncu kepler rc=$?"
PERF_N=1000000 PERF_E=100 ncu --set full --clock-control none --import-source on -k regex:'ephemeris_twobody' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_eph -f python tools/gpu_perf_eph.py > gpurun_out/${TAG}_ncu_eph.log 2>&1; echo "ncu eph rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lsq_quad_kernel' --launch-skip 1 -c 1 \
    -o gpurun_out/${TAG}_lsq -f python tools/gpu_perf_lsq.py 100000 > gpurun_out/${TAG}_ncu_lsq.log 2>&1; echo "ncu lsq rc=$?"
PERF_N=500000 ncu --set full --clock-control none --import-source on -k regex:'propagate_nbody_kernel' --launch-skip 1 -c 1 \
    -o gpurun_out/${TAG}_nbody -f python tools/gpu_perf_nbody.py > gpurun_out/${TAG}_ncu_nbody.log 2>&1; echo "ncu nbody rc=$?"
# text summaries on the box; the .ncu-rep files of the bulk kernels stay behind (gpurun merges at most 64 MiB back)
python tools/profile_report.py gpurun_out/${TAG}_phases.ncu-rep gpurun_out/${TAG}_ncu_phases.txt roots_kernel correct_kernel score_kernel > /dev/null 2>&1
python tools/ncu_latest.py gpurun_out/${TAG}_phases.ncu-rep gpurun_out/${TAG}_source_sha.txt c3_100k_x12 gpurun_out/${TAG}_ncu_latest.json > /dev/null 2>&1
for k in kepler:propagate_universal_kernel eph:ephemeris_twobody_kernel lsq:lsq_quad_kernel nbody:propagate_nbody_kernel; do
  python tools/profile_report.py gpurun_out/${TAG}_${k%%:*}.ncu-rep gpurun_out/${TAG}_ncu_${k%%:*}.txt ${k##*:} > /dev/null 2>&1
  [ "${KEEP_REPS:-0}" = "1" ] || rm -f gpurun_out/${TAG}_${k%%:*}.ncu-rep
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
fi
