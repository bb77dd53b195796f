set -x
nproc
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload batched --no-cpu-baseline > gpurun_out/r2D_bench_batched_8gpu.json 2> gpurun_out/r2D_bench_batched_8gpu.err ) 2> gpurun_out/r2D_time.txt; tail -2 gpurun_out/r2D_bench_batched_8gpu.err
