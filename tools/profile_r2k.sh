#!/bin/bash
# Round-2 closing evidence run for profiles/ (under gpurun, one GPU): the whole GPU test suite, bench lines of both arms taken OUTSIDE the
# profiler, the launch list of the bench command, ncu --set full of the final SYRK / Cholesky / K2 sequence / K1 kernels, KLTmain timing.
tag=${1:-r2k}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_ref.err
timeout 200 python tools/kltmain_bench.py > gpurun_out/${tag}_kltmain.txt 2>&1
timeout 200 python tools/c1_bench.py > gpurun_out/${tag}_c1.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_launches_bench.log 2>&1
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${tag}_$1 \
      python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_$1_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page details > gpurun_out/${tag}_ncu_$1_details.txt 2>&1
  rm -f gpurun_out/${tag}_$1.ncu-rep
}
cap syrk dsyrk_lower_sub 12
cap chol chol_dag 12
cap k2seq lk_seq_w15h 3
cap k1 pyrdown2_fused_tma 3
tail -2 gpurun_out/${tag}_kltmain.txt gpurun_out/${tag}_c1.txt
python - <<PY
import json
for n in ("bench","bench_reference_arm"):
    d=json.loads(open("gpurun_out/${tag}_%s.json"%n).read().strip().splitlines()[-1])
    print(n, round(d["value"],1), round(d["e2e"]["value"],1), d.get("clocks"), d.get("gpu_launches"))
PY
