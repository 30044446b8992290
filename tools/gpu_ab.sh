mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14 | tee gpurun_out/r02q_pytest.log
