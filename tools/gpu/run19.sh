set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi_context or sharded or sum" 2>&1 | tail -6 > gpurun_out/r2s_pytest_multi.log; cat gpurun_out/r2s_pytest_multi.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sum_and_resident or batch" 2>&1 | tail -4 >> gpurun_out/r2s_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_2gpu.json 2> gpurun_out/r2s_bench_2gpu.err; tail -3 gpurun_out/r2s_bench_2gpu.err
