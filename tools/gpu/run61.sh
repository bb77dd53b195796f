timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -1 gpurun_out/r2j_smoke.log
timeout 600 python bench.py > gpurun_out/r2j_bench_2p20_bls.json 2> gpurun_out/r2j_bench.err; tail -c 200 gpurun_out/r2j_bench.err
