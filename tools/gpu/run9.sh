set -x
timeout 600 python tools/exp_r2.py --sizes 12,14,16,18,20 --configs "base;fold_cluster=0" --tag foldc > gpurun_out/r2i_exp_foldc.jsonl 2>gpurun_out/r2i_exp.err
timeout 600 python tools/exp_r2.py --curve bn128 --sizes 16,18,20 --configs "base;fold_cluster=0" --tag foldc_bn > gpurun_out/r2i_exp_foldc_bn.jsonl 2>>gpurun_out/r2i_exp.err
timeout 600 python tools/exp_r2.py --curve bls12381_g2 --sizes 16,18 --configs "base;fold_cluster=0" --tag foldc_g2 > gpurun_out/r2i_exp_foldc_g2.jsonl 2>>gpurun_out/r2i_exp.err
tail -3 gpurun_out/r2i_exp.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_random or window_widths or edge_cases or groups_and_lanes or full_size or g2" 2>&1 | tail -4 > gpurun_out/r2i_pytest_core.log; cat gpurun_out/r2i_pytest_core.log
