set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2f_bench_8gpu_weak_2p20_each.json 2> gpurun_out/r2f_bench_8gpu.err; tail -2 gpurun_out/r2f_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload batched --no-cpu-baseline > gpurun_out/r2f_bench_batched_64x2p18_8gpu.json 2>> gpurun_out/r2f_bench_8gpu.err; tail -2 gpurun_out/r2f_bench_8gpu.err
