timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "lane_and_tail_variants" 2>&1 | tail -5 > gpurun_out/r2W_pytest.log; cat gpurun_out/r2W_pytest.log
