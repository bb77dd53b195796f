for rep in 1 2; do
python tools/exp_r2.py --lib tools/_bin/libold.so --tag old --sizes 16,18,20 --configs "base" --phases --iters 20 --reps 5 >> gpurun_out/r2P_exp.jsonl 2>>gpurun_out/r2P_exp.err
python tools/exp_r2.py --lib tools/_bin/libdev.so --tag dev --sizes 16,18,20 --configs "base;meta_upfront=0" --phases --iters 20 --reps 5 >> gpurun_out/r2P_exp.jsonl 2>>gpurun_out/r2P_exp.err
done
tail -3 gpurun_out/r2P_exp.err
