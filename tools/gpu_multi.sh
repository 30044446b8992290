mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/multi_${N}.json 2> gpurun_out/multi_${N}.err; echo "rc=$?"
tail -5 gpurun_out/multi_${N}.err; cut -c1-400 gpurun_out/multi_${N}.json
