set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_random or edge_cases or full_size" 2>&1 | tail -4 > gpurun_out/r2F_pytest_core.log; cat gpurun_out/r2F_pytest_core.log
timeout 300 python tools/trace_msm.py --log2n 16 --configs "base" --dump > gpurun_out/r2F_trace_2p16.txt 2>gpurun_out/r2F_trace.err
timeout 300 python tools/trace_msm.py --log2n 20 --configs "lanes=1" > gpurun_out/r2F_trace_2p20_l1.txt 2>>gpurun_out/r2F_trace.err
timeout 600 python tools/exp_r2.py --sizes 14,16,18,20 --configs "base" --tag ilp > gpurun_out/r2F_exp_ilp.jsonl 2>gpurun_out/r2F_exp.err
