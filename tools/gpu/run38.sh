for rep in 1 2; do
for l in old dev nosplit; do
python tools/exp_r2.py --lib tools/_bin/lib$l.so --tag $l --sizes 20 --configs "base" --phases --iters 20 --reps 5 >> gpurun_out/r2O_exp.jsonl 2>>gpurun_out/r2O_exp.err
done; done
tail -3 gpurun_out/r2O_exp.err
