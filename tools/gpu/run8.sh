set -x
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "batch_inverse or schedule" 2>&1 | tail -15 > gpurun_out/r2h_pytest_hooks.log; cat gpurun_out/r2h_pytest_hooks.log
timeout 600 python tools/exp_r2.py --sizes 14,16,18,20 --configs "base;issue_threads=0" --tag issue > gpurun_out/r2h_exp_issue.jsonl 2>gpurun_out/r2h_exp_issue.err
timeout 600 python tools/exp_r2.py --sizes 16,18,20 --windowed 0 --configs "base;issue_threads=0" --tag issue_win > gpurun_out/r2h_exp_issue_win.jsonl 2>>gpurun_out/r2h_exp_issue.err
timeout 600 python tools/exp_r2.py --curve bn128 --sizes 16,18,20 --configs "base;issue_threads=0" --tag issue_bn > gpurun_out/r2h_exp_issue_bn.jsonl 2>>gpurun_out/r2h_exp_issue.err
tail -3 gpurun_out/r2h_exp_issue.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
