set -x
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "g2_codecs" 2>&1 | tail -15 > gpurun_out/r2G_pytest.log; cat gpurun_out/r2G_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "codecs or ffjavascript" 2>&1 | tail -5 >> gpurun_out/r2G_pytest.log
