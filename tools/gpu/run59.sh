timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_digits|k_fold|k_accum_finish|k_inv_root|k_prod_fwd|k_prod_bwd|k_scan" -s 20 -c 16 -o /tmp/r2i_rest python tools/exp_r2.py --sizes 20 --configs "lanes=1" --iters 1 --reps 1 > gpurun_out/r2i_ncu_rest.log 2>&1
tail -2 gpurun_out/r2i_ncu_rest.log | cut -c1-200
ncu -i /tmp/r2i_rest.ncu-rep --page raw --csv > gpurun_out/r2i_rest_raw.csv 2>/dev/null
ls -la gpurun_out/r2i_rest_raw.csv
