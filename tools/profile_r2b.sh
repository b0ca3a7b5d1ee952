#!/bin/bash
# Final round-2 evidence run for profiles/ (under gpurun, one GPU):
#  (1) bench lines of both arms taken OUTSIDE the profiler,
#  (2) launch list (gpu__time_duration) of the bench command,
#  (3) ncu --set full of K1 (fused two-level TMA kernel), K2 batch, K2 sequence, the BA camera kernel, the SYRK, the Cholesky, K4.
tag=${1:-r2b}
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_launches_bench.log 2>&1
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${tag}_$1 \
      python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_$1_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page details > gpurun_out/${tag}_ncu_$1_details.txt 2>&1
}
cap k2 lk_track_w15h 2
cap k2seq lk_seq_w15h 3
cap k1 pyrdown2_fused_tma 3
cap k7cam bal_camera 12
cap syrk dsyrk_lower_sub 12
cap chol chol_dag 12
cap k4 knn2_hamming_tc 3
ls -la gpurun_out | tail -30
