#!/bin/bash
# Round-2 evidence run for profiles/ (under gpurun, one GPU):
#  (1) launch list (gpu__time_duration) of the bench command,
#  (2) ncu --set full of the K2 batch kernel (roofline leg), the K2 sequence kernel and K1,
#  (3) bench lines of both arms taken OUTSIDE the profiler.
tag=${1:-r2}
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lk_track_w15h -s 2 -c 1 -f -o gpurun_out/${tag}_k2 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_k2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lk_seq_w15h -s 3 -c 1 -f -o gpurun_out/${tag}_k2seq \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_k2seq_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyrdown -s 8 -c 2 -f -o gpurun_out/${tag}_k1 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_k1_ncu.log 2>&1
for k in k1 k2 k2seq; do ncu -i gpurun_out/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_ncu_${k}_details.txt 2>&1; done
ls -la gpurun_out | tail -14
