set -x
timeout 600 python tools/exp_r2.py --sizes 16,18,20,22 --configs "base;sort_groups=0;lanes=2;lanes=3;lanes=1" --phases --tag grouped > gpurun_out/r2f_exp.jsonl 2>gpurun_out/r2f_exp.err
tail -n 3 gpurun_out/r2f_exp.err
