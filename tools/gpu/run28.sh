for b in 8 16 32; do
timeout 600 python tools/exp_batch.py --batch $b --configs "base;batch_workers=4;batch_workers=4,batch_lanes=4;batch_workers=4,batch_lanes=2;batch_workers=2,batch_lanes=4;batch_workers=6;batch_workers=3,batch_lanes=2" >> gpurun_out/r2B_exp_batch_small.jsonl 2>> gpurun_out/r2B_exp.err
done
tail -2 gpurun_out/r2B_exp.err
