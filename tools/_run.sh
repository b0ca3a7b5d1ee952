python -m pytest tests -m gpu -x -q > gpurun_out/s3i_pytest.log 2>&1; tail -3 gpurun_out/s3i_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s3i_bench.json 2> gpurun_out/s3i_bench.err; tail -c 400 gpurun_out/s3i_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s3i_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['details']['stage_ms'],d['e2e']['value'],d['gpu_launches']); r=d['roofline']; print({k:r[k] for k in ('frac','ms_per_launch')}); print(r['k1_pyramid']); print(r.get('k8_schur_syrk')); print(r.get('k8_cholesky')); print(r.get('k4_match')); print(r.get('error'))"
