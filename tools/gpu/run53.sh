timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -1 gpurun_out/r2h_smoke.log
timeout 600 python bench.py > gpurun_out/r2h_bench_2p20_bls.json 2> gpurun_out/r2h_bench.err; tail -c 300 gpurun_out/r2h_bench.err
timeout 300 python bench.py --curve bn128 --no-cpu-baseline > gpurun_out/r2h_bench_2p20_bn254.json 2>> gpurun_out/r2h_bench.err
timeout 300 python bench.py --log2n 18 --no-cpu-baseline > gpurun_out/r2h_bench_2p18_bls.json 2>> gpurun_out/r2h_bench.err
timeout 300 python bench.py --log2n 16 --no-cpu-baseline > gpurun_out/r2h_bench_2p16_bls.json 2>> gpurun_out/r2h_bench.err
