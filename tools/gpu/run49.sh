B200MSM_LIB=$PWD/tools/_bin/libdev.so timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "batch_inverse" 2>&1 | tail -3 > gpurun_out/r2V_pytest.log; cat gpurun_out/r2V_pytest.log
for rep in 1 2; do
timeout 300 python tools/exp_r2.py --tag old --sizes 14,16,18,20 --configs "base" --iters 20 --reps 5 >> gpurun_out/r2V_exp.jsonl 2>>gpurun_out/r2V_exp.err
timeout 300 python tools/exp_r2.py --lib tools/_bin/libdev.so --tag new --sizes 14,16,18,20 --configs "base" --iters 20 --reps 5 >> gpurun_out/r2V_exp.jsonl 2>>gpurun_out/r2V_exp.err
done
tail -3 gpurun_out/r2V_exp.err
