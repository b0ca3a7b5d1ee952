python -m pytest tests/test_klt_gpu.py -x -q -k "pyramid" 2>&1 | tail -3
for t in 1 0; do
VEL_PYR_TMA=$t python bench.py --steps 5 --warmup 3 > gpurun_out/s3h_bench_$t.json 2> gpurun_out/s3h_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s3h_bench_$t.json').read().strip().splitlines()[-1]);print($t, d['value'],d['details']['stage_ms']['klt_pyramids_and_tracking'], d['roofline']['k1_pyramid'])"
done
