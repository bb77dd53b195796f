timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
timeout 600 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 300 gpurun_out/r2g_bench.err
for cv in bls12381 bn128; do
  timeout 600 python tools/sweep.py --curve $cv --sizes 10,12,14,16,18,20,22,24,26 --reps 3 --roofline 2>&1 | grep '^{' >> gpurun_out/r2g_sweep.jsonl
  timeout 600 python tools/sweep.py --curve $cv --sizes 10,12,14,16,18,20,22,24 --reps 3 --roofline --windowed 0 2>&1 | grep '^{' >> gpurun_out/r2g_sweep.jsonl
done
