mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 3 python tools/gpu_sanitize.py > gpurun_out/r02p_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02p_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 3 python tools/gpu_sanitize.py > gpurun_out/r02p_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02p_racecheck.log
