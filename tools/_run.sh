python -m pytest tests/test_klt_gpu.py tests/test_e2e.py tests/test_features_gpu.py tests/test_ransac_gpu.py -x -q 2>&1 | tail -4
python tools/kltmain_bench.py 2>&1 | tail -5
