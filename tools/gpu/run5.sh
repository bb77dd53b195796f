set -x
for v in lean5 lean6; do
timeout 300 python tools/exp_r2.py --sizes 18,20 --configs "base;bwd_lean=1;lanes=1;lanes=1,bwd_lean=1" --phases --tag $v --lib zprize-wasm-msm_b200/b200msm/variants/lib_$v.so > gpurun_out/r2e_exp_$v.jsonl 2>gpurun_out/r2e_exp_$v.err
tail -n 2 gpurun_out/r2e_exp_$v.err
done
