set -x
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "x_only" 2>&1 | tail -5 > gpurun_out/r2f_pytest_xonly.log; cat gpurun_out/r2f_pytest_xonly.log
timeout 600 python bench.py > gpurun_out/r2f_bench_2p20_bls.json 2> gpurun_out/r2f_bench.err
timeout 300 python bench.py --log2n 18 --no-cpu-baseline > gpurun_out/r2f_bench_2p18_bls.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --log2n 16 --no-cpu-baseline > gpurun_out/r2f_bench_2p16_bls.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --curve bn128 --no-cpu-baseline > gpurun_out/r2f_bench_2p20_bn254.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --curve bls12381_g2 --no-cpu-baseline > gpurun_out/r2f_bench_2p20_bls_g2.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --workload batched --no-cpu-baseline > gpurun_out/r2f_bench_batched_64x2p18.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --workload ntt --no-cpu-baseline > gpurun_out/r2f_bench_ntt_2p24.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference_arm_2p20.json 2>> gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_bench_py_2p20_bls.csv python bench.py --steps 2 --warmup 1 --no-window-table --no-cpu-baseline > gpurun_out/r2f_ncu_launches.log 2>&1
tail -2 gpurun_out/r2f_ncu_launches.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tree_bwd|k_tree_fwd|k_tree_meta|k_extract_x" -s 15 -c 7 -o /tmp/r2f_tree python tools/exp_r2.py --sizes 20 --configs "lanes=1" --iters 1 --reps 1 > gpurun_out/r2f_ncu_full.log 2>&1
tail -2 gpurun_out/r2f_ncu_full.log | cut -c1-300
ncu -i /tmp/r2f_tree.ncu-rep --page raw --csv > gpurun_out/r2f_tree_raw.csv 2>/dev/null
ls -la gpurun_out/r2f_tree_raw.csv
