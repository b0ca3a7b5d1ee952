#!/bin/bash
# ncu --set full of one K2 launch per kernel variant (args: `default` and/or `bytes`); raw pages land in gpurun_out/
for impl in "$@"; do
  if [ $impl = default ]; then unset VEL_LK_W15; else export VEL_LK_W15=$impl; fi
  python tools/lk_one.py 32 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_track_w15 --launch-skip 1 -c 1 -o gpurun_out/k2_$impl -f python tools/lk_one.py 32 2 > gpurun_out/k2_${impl}_ncu.log 2>&1
  ncu -i gpurun_out/k2_$impl.ncu-rep --page details > gpurun_out/k2_${impl}_details.txt 2>&1
  ncu -i gpurun_out/k2_$impl.ncu-rep --page source --csv > gpurun_out/k2_${impl}_source.csv 2>&1
done
