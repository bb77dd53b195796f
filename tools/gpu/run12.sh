set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_random or window_widths or edge_cases or groups_and_lanes or full_size or g2" 2>&1 | tail -4 > gpurun_out/r2l_pytest_core.log; cat gpurun_out/r2l_pytest_core.log
timeout 600 python tools/exp_r2.py --sizes 18,20 --configs "base;lanes=1;window_bits=15" --phases --tag dense2 > gpurun_out/r2l_exp_dense2.jsonl 2>gpurun_out/r2l_exp.err
timeout 600 python tools/exp_glv.py --sizes 16,18,20 --wb 0,14,15,16 > gpurun_out/r2l_exp_glv.jsonl 2>>gpurun_out/r2l_exp.err
tail -3 gpurun_out/r2l_exp.err
timeout 300 python tools/trace_msm.py --log2n 20 --configs "lanes=1" > gpurun_out/r2l_trace_2p20_l1.txt 2>gpurun_out/r2l_trace.err
