timeout 300 python tools/match_check.py 2>&1 | tail -8
timeout 300 python -m pytest tests/test_sfm_gpu.py -x -q -k match 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn2_hamming_tc|expand_pm" -s 12 -c 2 python tools/match_check.py 2>&1 | grep -E "duration" | tail -4
