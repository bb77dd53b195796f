timeout 600 python tools/exp_glv.py --sizes 14,16,18,20 --wb 0,13,14,15,16 > gpurun_out/r2m_exp_glv.jsonl 2>gpurun_out/r2m_exp.err
tail -3 gpurun_out/r2m_exp.err
