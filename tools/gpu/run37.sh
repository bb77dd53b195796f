for rep in 1 2; do
python tools/exp_r2.py --lib tools/_bin/libold.so --tag old --sizes 18,20 --configs "base" --phases --iters 20 --reps 5 >> gpurun_out/r2N_exp.jsonl 2>>gpurun_out/r2N_exp.err
python tools/exp_r2.py --lib tools/_bin/libdev.so --tag new --sizes 18,20 --configs "base;xonly=0" --phases --iters 20 --reps 5 >> gpurun_out/r2N_exp.jsonl 2>>gpurun_out/r2N_exp.err
done
tail -3 gpurun_out/r2N_exp.err
