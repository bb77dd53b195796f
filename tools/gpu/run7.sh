set -x
timeout 300 python tools/trace_msm.py --log2n 20 --configs "base" --dump --out gpurun_out/r2g_trace_2p20.json > gpurun_out/r2g_trace_2p20.txt 2>gpurun_out/r2g_trace.err
timeout 300 python tools/trace_msm.py --log2n 18 --configs "base" --dump --out gpurun_out/r2g_trace_2p18.json > gpurun_out/r2g_trace_2p18.txt 2>>gpurun_out/r2g_trace.err
timeout 300 python tools/trace_msm.py --log2n 16 --configs "base" --dump > gpurun_out/r2g_trace_2p16.txt 2>>gpurun_out/r2g_trace.err
tail -5 gpurun_out/r2g_trace.err
timeout 300 python tools/exp_r2.py --sizes 18,20 --configs "base;ba_k=12;ba_k=16;ba_k=32;pt_k=16;persist=740;persist=0" --tag k > gpurun_out/r2g_exp_k.jsonl 2>gpurun_out/r2g_exp_k.err
