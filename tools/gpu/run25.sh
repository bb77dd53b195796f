timeout 900 python tools/exp_r2.py --sizes 20 --configs "base;persist=0;lanes=8,persist=0;lanes=8,persist=0,sort_groups=0;lanes=6,persist=0;lanes=4,persist=0,sort_groups=0;lanes=8,persist=0,fold_cluster=0;lanes=8,persist=296;lanes=8,persist=148" --tag div > gpurun_out/r2y_exp_div.jsonl 2>gpurun_out/r2y_exp.err
timeout 900 python tools/exp_r2.py --sizes 18 --configs "base;persist=0;lanes=8,persist=0;lanes=8,persist=0,sort_groups=0;lanes=6,persist=0;lanes=8,persist=0,fold_cluster=0" --tag div >> gpurun_out/r2y_exp_div.jsonl 2>>gpurun_out/r2y_exp.err
tail -2 gpurun_out/r2y_exp.err
