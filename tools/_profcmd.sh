ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_2p16.csv python tools/sweep.py --sizes 16 --reps 1 > gpurun_out/ncu_2p16.log 2>&1
tail -2 gpurun_out/ncu_2p16.log | cut -c1-600
