set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2C_pytest.log; cat gpurun_out/r2C_pytest.log
timeout 600 python tools/exp_batch.py --configs "base" > gpurun_out/r2C_exp_batch.jsonl 2> gpurun_out/r2C_exp_batch.err; cat gpurun_out/r2C_exp_batch.jsonl
timeout 600 python tools/exp_batch.py --batch 8 --configs "base" >> gpurun_out/r2C_exp_batch.jsonl 2>> gpurun_out/r2C_exp_batch.err
timeout 900 python bench.py > gpurun_out/r2C_bench_1gpu.json 2> gpurun_out/r2C_bench_1gpu.err; tail -2 gpurun_out/r2C_bench_1gpu.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2C_smoke.log 2>&1; tail -1 gpurun_out/r2C_smoke.log
