python -m pytest tests/test_e2e.py -x -q 2>&1 | tail -2
python tools/c1_bench.py 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/s3p_bench.json 2> gpurun_out/s3p_bench.err; tail -c 300 gpurun_out/s3p_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s3p_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['details']['host_affinity'], d['cpu_baseline']['cores'])"
