timeout 300 python tools/dense_check.py 2>&1 | tail -10
timeout 300 python -m pytest tests/test_dense_gpu.py tests/test_sfm_gpu.py -x -q 2>&1 | tail -2
timeout 120 python tools/ba_c3_iter.py 2>&1 | tail -4
