python -m pytest tests/test_sequence_gpu.py -x -q 2>&1 | tail -3
python tools/e2e_probe.py 2>&1 | head -4
python bench.py --steps 10 --warmup 3 > gpurun_out/s3k_bench.json 2> gpurun_out/s3k_bench.err; tail -c 300 gpurun_out/s3k_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s3k_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'])"
