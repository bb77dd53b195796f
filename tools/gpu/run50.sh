timeout 300 python bench.py --workload ntt --no-cpu-baseline > gpurun_out/r2g_bench_ntt_2p24.json 2> gpurun_out/r2g_bench_ntt.err; tail -2 gpurun_out/r2g_bench_ntt.err
