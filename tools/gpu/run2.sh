set -x
# correctness first, with a hard timeout (the fused round kernel spins on flags)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_random or window_widths or edge_cases or groups_and_lanes or full_size" 2>&1 | tail -8 > gpurun_out/r2b_pytest_core.log; cat gpurun_out/r2b_pytest_core.log
timeout 600 python tools/exp_r2.py --sizes 14,16,18,20 --configs "base;lanes=1;lanes=2;fused=0;fused=0,lanes=1" --phases --tag fused > gpurun_out/r2b_exp_fused.jsonl 2>gpurun_out/r2b_exp_fused.err
tail -n 3 gpurun_out/r2b_exp_fused.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_pytest_all.log; cat gpurun_out/r2b_pytest_all.log
