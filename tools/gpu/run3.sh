set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tree_round -s 4 -c 3 -o gpurun_out/r2c_round python tools/exp_r2.py --sizes 20 --configs "lanes=1" --iters 1 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
tail -5 gpurun_out/r2c_ncu.log
ncu -i gpurun_out/r2c_round.ncu-rep --page raw --csv > gpurun_out/r2c_round_raw.csv 2>/dev/null
ls -la gpurun_out/
