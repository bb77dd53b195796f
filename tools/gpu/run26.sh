set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or sum_and_resident or window_table" 2>&1 | tail -4 > gpurun_out/r2z_pytest_batch.log; cat gpurun_out/r2z_pytest_batch.log
timeout 600 python tools/exp_batch.py --configs "base;batch_workers=4;batch_lanes=2" > gpurun_out/r2z_exp_batch.jsonl 2> gpurun_out/r2z_exp_batch.err; tail -3 gpurun_out/r2z_exp_batch.err
timeout 600 python bench.py --workload batched --no-cpu-baseline > gpurun_out/r2z_bench_batched_64x2p18.json 2>gpurun_out/r2z_bench.err; tail -3 gpurun_out/r2z_bench.err
