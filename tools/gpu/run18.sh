set -x
for wb in 18 19 20 21; do
timeout 600 python tools/exp_r2.py --sizes 20 --windowed $wb --configs "base;subslots=8;subslots=2" --phases --tag win$wb >> gpurun_out/r2r_exp_win.jsonl 2>>gpurun_out/r2r_exp.err
done
timeout 600 python tools/exp_r2.py --sizes 18 --windowed 17 --configs "base" --tag win17 >> gpurun_out/r2r_exp_win.jsonl 2>>gpurun_out/r2r_exp.err
timeout 600 python tools/exp_r2.py --sizes 18 --windowed 18 --configs "base" --tag win18 >> gpurun_out/r2r_exp_win.jsonl 2>>gpurun_out/r2r_exp.err
timeout 600 python tools/exp_r2.py --sizes 18 --windowed 16 --configs "base" --tag win16 >> gpurun_out/r2r_exp_win.jsonl 2>>gpurun_out/r2r_exp.err
tail -3 gpurun_out/r2r_exp.err
