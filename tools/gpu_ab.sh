mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02i_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02i_bench.json'));print(d['value'],d['e2e']['value'],d['e2e']['ms_per_step']);print(d['e2e']['seeded'])"
