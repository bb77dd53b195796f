set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 600 python tools/exp_r2.py --sizes 14,16,18,20 --configs "base;lanes=1" --phases --tag main > gpurun_out/r2a_exp_main.jsonl 2>gpurun_out/r2a_exp_main.err
timeout 300 python tools/exp_r2.py --sizes 18,20 --configs "base;l2_persist=50;l2_persist=100;lanes=1;lanes=1,l2_persist=100" --phases --tag l2 > gpurun_out/r2a_exp_l2.jsonl 2>gpurun_out/r2a_exp_l2.err
timeout 300 python tools/exp_r2.py --sizes 18,20 --configs "base;lanes=1" --phases --tag mul2 --lib zprize-wasm-msm_b200/b200msm/variants/libb200msm_mul2.so > gpurun_out/r2a_exp_mul2.jsonl 2>gpurun_out/r2a_exp_mul2.err
python - <<'PY' > gpurun_out/r2a_l2info.txt 2>&1
import sys; sys.path[:0]=['.','zprize-wasm-msm_b200']
import b200msm
e=b200msm.Engine(0); print('l2_persist_max', e.counter('l2_persist_max_bytes'), 'window_max', e.counter('l2_window_max_bytes'))
PY
tail -3 gpurun_out/*.err
