#!/bin/bash
# Evidence run for profiles/: (1) launch list of the bench command, (2) ncu --set full of the dominant kernel (K2) and of
# K1 inside the same command.  Usage (under gpurun): bash tools/profile_bench.sh <tag>
tag=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lk_track -s 3 -c 1 -f -o gpurun_out/${tag}_k2 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_k2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyrdown -s 6 -c 2 -f -o gpurun_out/${tag}_k1 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_k1_ncu.log 2>&1
for k in k1 k2; do ncu -i gpurun_out/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_ncu_${k}_details.txt 2>&1; done
ls -la gpurun_out | tail -12
