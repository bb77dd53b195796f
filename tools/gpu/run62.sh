timeout 200 python bench.py --log2n 18 --no-cpu-baseline > gpurun_out/r2j_bench_2p18_bls.json 2> gpurun_out/r2j_bench2.err
timeout 200 python bench.py --log2n 16 --no-cpu-baseline > gpurun_out/r2j_bench_2p16_bls.json 2>> gpurun_out/r2j_bench2.err
timeout 200 python bench.py --curve bn128 --no-cpu-baseline > gpurun_out/r2j_bench_2p20_bn254.json 2>> gpurun_out/r2j_bench2.err
timeout 200 python bench.py --workload batched --no-cpu-baseline > gpurun_out/r2j_bench_batched_64x2p18.json 2>> gpurun_out/r2j_bench2.err
tail -2 gpurun_out/r2j_bench2.err
