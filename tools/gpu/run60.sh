for rep in 1 2; do
timeout 300 python tools/exp_r2.py --tag base --sizes 18,20,22 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2b_exp.jsonl 2>>gpurun_out/r2b_exp.err
timeout 300 python tools/exp_r2.py --lib tools/_bin/libv10.so --tag fwd5 --sizes 18,20,22 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2b_exp.jsonl 2>>gpurun_out/r2b_exp.err
done
tail -3 gpurun_out/r2b_exp.err
