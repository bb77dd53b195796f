timeout 120 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 16 --configs "base;fused_round=2,lanes=1" > gpurun_out/r2R_exp.jsonl 2>gpurun_out/r2R_exp.err
tail -2 gpurun_out/r2R_exp.err; cat gpurun_out/r2R_exp.jsonl
timeout 300 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 16,18,20 --configs "base;fused_round=2;fused_round=2,lanes=2;fused_round=2,lanes=1;fused_round=2,lanes=2,fused_kmax=8;fused_round=2,lanes=2,fused_kmax=32;fused_round=2,lanes=2,fused_tiles=444;fused_round=2,lanes=2,fused_tiles=1776" >> gpurun_out/r2R_exp.jsonl 2>>gpurun_out/r2R_exp.err
timeout 200 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 18,20 --configs "base;fused_round=2,lanes=2" --phases --curve bn128 >> gpurun_out/r2R_exp.jsonl 2>>gpurun_out/r2R_exp.err
timeout 200 python tools/exp_r2.py --lib tools/_bin/libexp.so --sizes 20 --configs "base;fused_round=2,lanes=2" --phases >> gpurun_out/r2R_exp.jsonl 2>>gpurun_out/r2R_exp.err
tail -3 gpurun_out/r2R_exp.err
