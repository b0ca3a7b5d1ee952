python -m pytest tests/test_sequence_gpu.py -x -q 2>&1 | tail -3
VEL_LK_SEQ=single python -m pytest tests/test_sequence_gpu.py -x -q 2>&1 | tail -2
for m in dual single; do
VEL_LK_SEQ=$m python bench.py --steps 5 --warmup 3 > gpurun_out/s3m_bench_$m.json 2> gpurun_out/s3m_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s3m_bench_$m.json').read().strip().splitlines()[-1]);print('$m', d['value'],d['ms_per_step'],d['details']['stage_ms']['klt_pyramids_and_tracking'], d['result']['speed_kmh_mean'])"
done
