#!/bin/bash
# 2-GPU box: the group tests on two real devices, then the bench line under torchrun (c4 strong scaling + group leg)
mkdir -p gpurun_out
TAG=${1:-r2f}
nvidia-smi -L | tee gpurun_out/${TAG}_gpus.txt
python -m pytest tests/test_group_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_group.log 2>&1; echo "group tests rc=$?"; tail -3 gpurun_out/${TAG}_pytest_group.log
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 \
   > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_n$N.err; cut -c1-300 gpurun_out/${TAG}_bench_n$N.json
