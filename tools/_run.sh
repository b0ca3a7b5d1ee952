for h in 16 24 32; do
VEL_PYR_H2=$h python bench.py --steps 5 --warmup 3 > gpurun_out/s3g_bench_$h.json 2> gpurun_out/s3g_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s3g_bench_$h.json').read().strip().splitlines()[-1]);print($h, d['value'],d['details']['stage_ms']['klt_pyramids_and_tracking'], d['roofline']['k1_pyramid'])"
done
VEL_PYR_H2=32 python -m pytest tests/test_klt_gpu.py -x -q -k "pyramid" 2>&1 | tail -2
VEL_PYR_H2=24 python -m pytest tests/test_klt_gpu.py -x -q -k "pyramid" 2>&1 | tail -2
