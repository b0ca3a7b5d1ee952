timeout 300 python -m pytest tests/test_sfm_gpu.py tests/test_dense_gpu.py tests/test_sequence_gpu.py -x -q 2>&1 | tail -2
timeout 120 python tools/ba_c3_iter.py 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bal_camera" -s 40 -c 2 python tools/ba_c3_iter.py 2>&1 | grep -E "bal_camera|duration" | tail -4
