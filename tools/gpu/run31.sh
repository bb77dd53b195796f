set -x
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "skewed" 2>&1 | tail -12 > gpurun_out/r2E_pytest2.log; cat gpurun_out/r2E_pytest2.log
