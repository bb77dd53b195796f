set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_random or window_widths or edge_cases or groups_and_lanes or full_size or g2" 2>&1 | tail -4 > gpurun_out/r2n_pytest_core.log; cat gpurun_out/r2n_pytest_core.log
timeout 600 python tools/exp_r2.py --sizes 16,18,20,22 --configs "base;lanes=1" --phases --tag dense3 > gpurun_out/r2n_exp_dense3.jsonl 2>gpurun_out/r2n_exp.err
timeout 600 python tools/exp_r2.py --sizes 18,20 --windowed 0 --configs "base" --tag dense3_win > gpurun_out/r2n_exp_dense3_win.jsonl 2>>gpurun_out/r2n_exp.err
tail -3 gpurun_out/r2n_exp.err
timeout 300 python tools/trace_msm.py --log2n 20 --configs "lanes=1" > gpurun_out/r2n_trace_2p20_l1.txt 2>gpurun_out/r2n_trace.err
