python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 19,20,22 --configs "xonly=0;base" --phases > gpurun_out/r2L_exp.jsonl 2>gpurun_out/r2L_exp.err
python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 20 --configs "xonly=0;base" --curve bn128 >> gpurun_out/r2L_exp.jsonl 2>>gpurun_out/r2L_exp.err
python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 20,24 --configs "xonly=0;base;xonly=0;base" --iters 20 --reps 5 >> gpurun_out/r2L_exp.jsonl 2>>gpurun_out/r2L_exp.err
tail -3 gpurun_out/r2L_exp.err
