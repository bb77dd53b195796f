timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2M_pytest.log 2>&1; tail -3 gpurun_out/r2M_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2M_smoke.log 2>&1; tail -1 gpurun_out/r2M_smoke.log
timeout 600 python bench.py > gpurun_out/r2M_bench.json 2> gpurun_out/r2M_bench.err; tail -c 300 gpurun_out/r2M_bench.err
