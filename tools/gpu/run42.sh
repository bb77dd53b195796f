set -x
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "all_visible_gpus or two_ranks_nccl" 2>&1 | tail -5 > gpurun_out/r2f_pytest_2gpu.log; cat gpurun_out/r2f_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_2gpu_weak_2p20_each.json 2> gpurun_out/r2f_bench_2gpu.err; tail -2 gpurun_out/r2f_bench_2gpu.err
