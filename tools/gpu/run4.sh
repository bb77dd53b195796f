set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_random or window_widths or edge_cases or groups_and_lanes or full_size" 2>&1 | tail -4 > gpurun_out/r2d_pytest_core.log; cat gpurun_out/r2d_pytest_core.log
timeout 600 python tools/exp_r2.py --sizes 16,18,20 --configs "base;lanes=1;lanes=2;fused=0" --phases --tag rootcta > gpurun_out/r2d_exp.jsonl 2>gpurun_out/r2d_exp.err
tail -n 3 gpurun_out/r2d_exp.err
timeout 600 ncu --set full --clock-control none -k regex:k_tree_round -s 5 -c 2 -o /tmp/r2d_round python tools/exp_r2.py --sizes 20 --configs "lanes=1" --iters 1 --reps 1 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
ncu -i /tmp/r2d_round.ncu-rep --page raw --csv > gpurun_out/r2d_round_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/r2d_round.ncu-rep
