timeout 300 python tools/exp_r2.py --sizes 18,20,22 --configs "base;persist_fwd=740;persist_fwd=444;base;persist_fwd=740" --iters 20 --reps 5 > gpurun_out/r2c_exp.jsonl 2>gpurun_out/r2c_exp.err
tail -2 gpurun_out/r2c_exp.err
