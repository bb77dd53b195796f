ncu --set full --clock-control none --import-source on -k regex:"k_ntt_tile|k_ntt_stage4|k_ntt_bitrev" -s 3 -c 4 -o gpurun_out/prof_r1f_ntt python tools/ntt_run.py --log2n 24 --reps 1 > gpurun_out/ncu_r1f.log 2>&1
ncu --set full --clock-control none -k regex:"k_tree_bwd|k_tree_fwd|k_fold$" -s 12 -c 3 -o /tmp/prof_r1f_g2 python tools/sweep.py --curve bls12381_g2 --sizes 20 --reps 1 --opt lanes=1 > gpurun_out/ncu_r1f_g2.log 2>&1
ncu -i /tmp/prof_r1f_g2.ncu-rep --page raw --csv > gpurun_out/prof_r1f_g2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
