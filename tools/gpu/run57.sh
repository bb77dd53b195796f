for rep in 1 2; do
timeout 300 python tools/exp_r2.py --tag base --sizes 16,18,20 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2a_exp.jsonl 2>>gpurun_out/r2a_exp.err
timeout 300 python tools/exp_r2.py --lib tools/_bin/libv8.so --tag reorder --sizes 16,18,20 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2a_exp.jsonl 2>>gpurun_out/r2a_exp.err
timeout 300 python tools/exp_r2.py --lib tools/_bin/libv9.so --tag reorder_4cta --sizes 16,18,20 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2a_exp.jsonl 2>>gpurun_out/r2a_exp.err
done
timeout 300 python tools/exp_r2.py --tag base --curve bn128 --sizes 18,20 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2a_exp.jsonl 2>>gpurun_out/r2a_exp.err
timeout 300 python tools/exp_r2.py --lib tools/_bin/libv9.so --tag reorder_4cta --curve bn128 --sizes 18,20 --configs "base" --iters 20 --reps 5 --phases >> gpurun_out/r2a_exp.jsonl 2>>gpurun_out/r2a_exp.err
tail -3 gpurun_out/r2a_exp.err
