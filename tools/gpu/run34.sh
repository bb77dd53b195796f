export B200MSM_LIB=$PWD/tools/_bin/libdev.so
python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 20 --configs "block_tree=0;block_tree=0,tree_rounds=4;block_tree=0,tree_rounds=3;block_tree=0,group_plan=102533;block_tree=0,group_plan=168067;block_tree=0,group_plan=34672772;block_tree=0,group_plan=35468184675;block_tree=0,group_plan=70936234050" > gpurun_out/r2K_exp_single.jsonl 2>gpurun_out/r2K_exp.err
python tools/exp_r2.py --lib tools/_bin/libdev.so --sizes 22 --configs "block_tree=0;block_tree=0,tree_rounds=4;block_tree=0,tree_rounds=5" >> gpurun_out/r2K_exp_single.jsonl 2>>gpurun_out/r2K_exp.err
python tools/exp_batch.py --configs "batch_workers=8,block_tree=0;batch_workers=8;batch_workers=8,fused_round=1;batch_workers=8,fused_round=1,fused_kmax=8;batch_workers=12,fused_round=1" > gpurun_out/r2K_exp.jsonl 2>>gpurun_out/r2K_exp.err
python tools/exp_batch.py --log2n 20 --batch 16 --configs "batch_workers=8,block_tree=0;batch_workers=8;batch_workers=8,fused_round=1;batch_workers=8,fused_round=1,fused_kmax=32" >> gpurun_out/r2K_exp.jsonl 2>>gpurun_out/r2K_exp.err
python tools/exp_batch.py --log2n 16 --batch 64 --configs "batch_workers=8,block_tree=0;batch_workers=8;batch_workers=8,fused_round=1" >> gpurun_out/r2K_exp.jsonl 2>>gpurun_out/r2K_exp.err
tail -3 gpurun_out/r2K_exp.err
