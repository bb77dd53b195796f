set -x
bash tools/_sweepcmd.sh > gpurun_out/r2w_sweep_1gpu.jsonl 2> gpurun_out/r2w_sweep.err
tail -2 gpurun_out/r2w_sweep.err
for cv in bls12381_g2 bn128_g2; do python tools/sweep.py --curve $cv --sizes 14,16,18,20 --reps 3 --roofline 2>&1 | grep '^{' >> gpurun_out/r2w_sweep_g2.jsonl; done
timeout 600 python bench.py --log2n 18 --no-cpu-baseline > gpurun_out/r2w_bench_2p18_bls.json 2>gpurun_out/r2w_bench.err
timeout 600 python bench.py --log2n 16 --no-cpu-baseline > gpurun_out/r2w_bench_2p16_bls.json 2>>gpurun_out/r2w_bench.err
timeout 600 python bench.py --curve bn128 --no-cpu-baseline > gpurun_out/r2w_bench_2p20_bn254.json 2>>gpurun_out/r2w_bench.err
timeout 600 python bench.py --curve bls12381_g2 --no-cpu-baseline > gpurun_out/r2w_bench_2p20_bls_g2.json 2>>gpurun_out/r2w_bench.err
timeout 600 python bench.py --workload batched --no-cpu-baseline > gpurun_out/r2w_bench_batched_64x2p18.json 2>>gpurun_out/r2w_bench.err
timeout 600 python bench.py --workload ntt --no-cpu-baseline > gpurun_out/r2w_bench_ntt_2p24.json 2>>gpurun_out/r2w_bench.err
tail -3 gpurun_out/r2w_bench.err
