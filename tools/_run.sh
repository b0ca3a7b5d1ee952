python tools/ba_c3_iter.py > gpurun_out/s3b_ba_iter.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s3b_ba_launches.csv python tools/ba_c3_iter.py > /dev/null 2>&1
python -m pytest tests/test_sfm_gpu.py tests/test_dense_gpu.py tests/test_sequence_gpu.py -x -q > gpurun_out/s3b_pytest.log 2>&1
tail -5 gpurun_out/s3b_pytest.log; cat gpurun_out/s3b_ba_iter.log
