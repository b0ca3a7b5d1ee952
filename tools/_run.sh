timeout 600 python -m pytest tests/test_sfm_gpu.py tests/test_dense_gpu.py tests/test_sequence_gpu.py -x -q 2>&1 | tail -2
timeout 300 python tools/dense_large.py 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['details']['stage_ms'],d['e2e']['value'],d['gpu_launches']); r=d['roofline']; print(r['frac'], r['k1_pyramid']['frac'], r['k8_schur_syrk']['frac'], r['k8_schur_syrk']['ms_per_launch'], r['k8_cholesky']['ms_per_launch'], r['k4_match']['us_per_frame_pair'])"
