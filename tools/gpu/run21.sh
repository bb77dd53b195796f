set -x
timeout 900 python tools/exp_r2.py --sizes 20 --configs "base;persist=444,persist_fwd=148;persist=444,persist_fwd=296;persist=444,persist_fwd=444;persist=518,persist_fwd=148;persist=592,persist_fwd=296;persist=592,persist_fwd=148;lanes=6;lanes=8;lanes=5" --tag ps > gpurun_out/r2u_exp_ps.jsonl 2>gpurun_out/r2u_exp.err
tail -3 gpurun_out/r2u_exp.err
