timeout 120 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 16 --configs "base;fused_round=3,lanes=1" > gpurun_out/r2S_exp.jsonl 2>gpurun_out/r2S_exp.err
tail -2 gpurun_out/r2S_exp.err; cat gpurun_out/r2S_exp.jsonl
timeout 300 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 18,20 --configs "base;fused_round=3,lanes=2;fused_round=3,lanes=1;fused_round=3,lanes=2,fused_kmax=12;fused_round=3,lanes=2,fused_kmax=24;fused_round=3,lanes=3,fused_kmax=12;fused_round=2,lanes=2,fused_kmax=12" >> gpurun_out/r2S_exp.jsonl 2>>gpurun_out/r2S_exp.err
timeout 200 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 20 --configs "base;fused_round=3,lanes=2,fused_kmax=12;fused_round=2,lanes=2,fused_kmax=12" --phases >> gpurun_out/r2S_exp.jsonl 2>>gpurun_out/r2S_exp.err
tail -3 gpurun_out/r2S_exp.err
