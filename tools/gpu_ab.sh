mkdir -p gpurun_out
TAG=${1:-ab}
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print(d['value'],d['e2e']['value']);print(d['ephemeris']);print(d['kepler'])"
