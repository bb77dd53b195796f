python tools/exp_r2.py --sizes 18,20,22 --configs "base;ba_k=6;ba_k=7;ba_k=10;ba_k=12;ba_k=14;ba_k=16;persist=444;persist=740;persist=0" --iters 20 --reps 3 > gpurun_out/r2Q_exp.jsonl 2>gpurun_out/r2Q_exp.err
tail -3 gpurun_out/r2Q_exp.err
