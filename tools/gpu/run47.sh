timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tree_round_pipe" -s 6 -c 2 -o /tmp/r2T_pipe python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 20 --configs "fused_round=2,lanes=1" --iters 1 --reps 1 > gpurun_out/r2T_ncu.log 2>&1
tail -3 gpurun_out/r2T_ncu.log | cut -c1-200
ncu -i /tmp/r2T_pipe.ncu-rep --page raw --csv > gpurun_out/r2T_pipe_raw.csv 2>/dev/null
ncu -i /tmp/r2T_pipe.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Issue Slots\|Registers\|Local\|Achieved Occupancy\|Theoretical Occ\|Executed Ipc\|Pipe\b" | head -60 > gpurun_out/r2T_pipe_details.txt
ls -la gpurun_out/r2T_*
