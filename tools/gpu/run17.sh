set -x
timeout 900 python tools/exp_r2.py --sizes 18,20,22 --configs "base;persist_fwd=888;persist_fwd=740;persist=740,persist_fwd=888;persist=592,persist_fwd=0" --tag pf > gpurun_out/r2q_exp_pf.jsonl 2>gpurun_out/r2q_exp.err
timeout 900 python tools/exp_r2.py --sizes 17 --configs "base;tree_rounds=1;tree_rounds=2;tree_rounds=3" --tag r17 > gpurun_out/r2q_exp_r17.jsonl 2>>gpurun_out/r2q_exp.err
timeout 900 python tools/exp_r2.py --sizes 19 --configs "base;tree_rounds=3;tree_rounds=4;tree_rounds=5" --tag r19 > gpurun_out/r2q_exp_r19.jsonl 2>>gpurun_out/r2q_exp.err
timeout 900 python tools/exp_r2.py --curve bn128 --sizes 16,18,20 --configs "base;persist_fwd=888" --tag pfbn > gpurun_out/r2q_exp_pfbn.jsonl 2>>gpurun_out/r2q_exp.err
tail -3 gpurun_out/r2q_exp.err
