set -x
timeout 300 python tools/trace_msm.py --log2n 20 --configs "lanes=1" --dump > gpurun_out/r2k_trace_2p20_l1.txt 2>gpurun_out/r2k_trace.err
timeout 300 python tools/trace_msm.py --log2n 18 --configs "lanes=1" --dump > gpurun_out/r2k_trace_2p18_l1.txt 2>>gpurun_out/r2k_trace.err
