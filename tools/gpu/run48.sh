timeout 120 python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 14 --configs "base;fold_cluster=2" > gpurun_out/r2U_exp.jsonl 2>gpurun_out/r2U_exp.err
tail -2 gpurun_out/r2U_exp.err; cat gpurun_out/r2U_exp.jsonl
timeout 300 python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 12,16,18,20 --configs "base;fold_cluster=2;base;fold_cluster=2" --iters 20 --reps 5 >> gpurun_out/r2U_exp.jsonl 2>>gpurun_out/r2U_exp.err
timeout 300 python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 16,18,20 --configs "base;fold_cluster=2" --curve bn128 --iters 20 --reps 5 >> gpurun_out/r2U_exp.jsonl 2>>gpurun_out/r2U_exp.err
B200MSM_LIB=$PWD/tools/_bin/libdev.so timeout 200 python tools/trace_msm.py --log2n 16 --configs "fold_cluster=2" > gpurun_out/r2U_trace_2p16.txt 2>>gpurun_out/r2U_exp.err
tail -3 gpurun_out/r2U_exp.err
