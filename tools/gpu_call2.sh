#!/bin/bash
# round 2, call 2: all GPU tests (incl. the full parity sweep), correct_kernel block-size A/B, the new bench line,
# ncu --set full captures of every hot kernel at HEAD, the launch list of the bench command
mkdir -p gpurun_out
TAG=r2b
python -c "import bench; print(bench.kernel_source_sha())" > gpurun_out/${TAG}_source_sha.txt
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
for v in default outfit_b200/variants/lib_bps4.so outfit_b200/variants/lib_ct64_bps10.so outfit_b200/variants/lib_ct32_bps20.so; do
  if [ $v = default ]; then unset OUTFIT_B200_LIB; else export OUTFIT_B200_LIB=$v; fi
  PERF_T=100000 OUTFIT_B200_STREAMS=1 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "LIB=|phases" | tee -a gpurun_out/${TAG}_correct_ab.log
  PERF_T=100000 PERF_PARITY=0 python tools/gpu_perf.py 2>&1 | grep -E "LIB=" | tee -a gpurun_out/${TAG}_correct_ab.log
done
unset OUTFIT_B200_LIB
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-800 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference arm rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
OUTFIT_B200_STREAMS=1 PERF_T=100000 PERF_PARITY=0 ncu --set full --clock-control none --import-source on -k regex:'roots_kernel|correct_kernel|score_kernel' \
    --launch-skip 3 -c 3 -o gpurun_out/${TAG}_phases -f python tools/gpu_perf.py > gpurun_out/${TAG}_ncu_phases.log 2>&1; echo "ncu phases rc=$?"
PERF_N=10000000 ncu --set full --clock-control none --import-source on -k regex:'propagate_universal_kernel' --launch-skip 2 -c 1 \
    -o gpurun_out/${TAG}_kepler -f python tools/gpu_perf_kepler.py > gpurun_out/${TAG}_ncu_kepler.log 2>&1; echo "ncu kepler rc=$?"
PERF_N=1000000 PERF_E=100 ncu --set full --clock-control none --import-source on -k regex:'ephemeris_' --launch-skip 2 -c 2 \
    -o gpurun_out/${TAG}_eph -f python tools/gpu_perf_eph.py > gpurun_out/${TAG}_ncu_eph.log 2>&1; echo "ncu eph rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lsq_kernel' --launch-skip 1 -c 1 \
    -o gpurun_out/${TAG}_lsq -f python tools/gpu_perf_lsq.py 100000 > gpurun_out/${TAG}_ncu_lsq.log 2>&1; echo "ncu lsq rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out | tail -20
