set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 1200 python bench.py > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench_1gpu.err; tail -2 gpurun_out/r2p_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err; tail -2 gpurun_out/r2p_bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tree_bwd|k_tree_fwd|k_tree_meta" -s 15 -c 6 -o /tmp/r2p_tree python tools/exp_r2.py --sizes 20 --configs "lanes=1" --iters 1 --reps 1 > gpurun_out/r2p_ncu_tree.log 2>&1; tail -2 gpurun_out/r2p_ncu_tree.log
ncu -i /tmp/r2p_tree.ncu-rep --page raw --csv > gpurun_out/r2p_tree_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2p_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-window-table --no-cpu-baseline > gpurun_out/r2p_ncu_bench.log 2>&1; tail -2 gpurun_out/r2p_ncu_bench.log
ls -la /tmp/r2p_tree.ncu-rep gpurun_out/
