set -x
timeout 900 python tools/exp_r2.py --sizes 20 --configs "base;group_small=40;group_small=100;groups=5;groups=5,group_small=30;groups=6,group_small=50;groups=8;lanes=3,groups=6;lanes=5,groups=5" --tag grp > gpurun_out/r2o_exp_grp20.jsonl 2>gpurun_out/r2o_exp.err
timeout 900 python tools/exp_r2.py --sizes 18 --configs "base;tree_rounds=4;tree_rounds=5;group_small=40;group_small=100;groups=5,group_small=30;groups=8;lanes=3;lanes=2" --tag grp > gpurun_out/r2o_exp_grp18.jsonl 2>>gpurun_out/r2o_exp.err
timeout 900 python tools/exp_r2.py --sizes 16 --configs "base;tree_rounds=2;tree_rounds=3;lanes=2;lanes=1;window_bits=12;window_bits=14" --tag grp > gpurun_out/r2o_exp_grp16.jsonl 2>>gpurun_out/r2o_exp.err
tail -3 gpurun_out/r2o_exp.err
