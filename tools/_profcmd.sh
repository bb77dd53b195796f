set -x
python tools/sweep.py --sizes 20 --reps 2 --opt lanes=1 > gpurun_out/plain_r1e.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1e.csv python tools/sweep.py --sizes 20 --reps 1 --opt lanes=1 > gpurun_out/ncu_r1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tree_bwd|k_tree_fwd|k_fold$|k_tree_meta|k_digits" -s 40 -c 14 -o gpurun_out/prof_r1e python tools/sweep.py --sizes 20 --reps 1 --opt lanes=1 > gpurun_out/ncu_r1e_full.log 2>&1
ls -la gpurun_out/
