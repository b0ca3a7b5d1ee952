python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bal_camera_kernel -s 12 -c 1 -f -o gpurun_out/r2c_k7cam python bench.py --steps 1 --warmup 3 > gpurun_out/r2c_k7cam_ncu.log 2>&1
ncu -i gpurun_out/r2c_k7cam.ncu-rep --page details > gpurun_out/r2c_ncu_k7cam_details.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2c_launches_bench.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['details']['stage_ms'],d['e2e']['value'],d['gpu_launches']); r=d['roofline']; print(r['frac'], r['k1_pyramid']['frac'], r['k8_schur_syrk']['frac'], r['k8_schur_syrk']['ms_per_launch'], r['k8_cholesky']['ms_per_launch'], r['k4_match']['us_per_frame_pair'])
d=json.loads(open('gpurun_out/r2c_bench_reference_arm.json').read().strip().splitlines()[-1]);print(d['value'])"
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots" gpurun_out/r2c_ncu_k7cam_details.txt | head
