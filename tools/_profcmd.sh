set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/f_bls.json 2> gpurun_out/f.err || exit 1
python bench.py --steps 10 --warmup 3 --log2n 18 > gpurun_out/f_bls18.json 2>> gpurun_out/f.err
python bench.py --steps 10 --warmup 3 --curve bn128 > gpurun_out/f_bn.json 2>> gpurun_out/f.err
python bench.py --steps 10 --warmup 3 --curve bls12381_g2 > gpurun_out/f_blsg2.json 2>> gpurun_out/f.err
python bench.py --steps 10 --warmup 3 --curve bn128_g2 > gpurun_out/f_bng2.json 2>> gpurun_out/f.err
python bench.py --workload batched --steps 5 --warmup 3 > gpurun_out/f_batched.json 2>> gpurun_out/f.err
python bench.py --workload ntt --steps 10 --warmup 3 > gpurun_out/f_ntt.json 2>> gpurun_out/f.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref.json 2>> gpurun_out/f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-window-table > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/f.err
for f in f_bls f_bls18 f_bn f_blsg2 f_bng2 f_batched; do python -c "
import json,sys; d=json.load(open('gpurun_out/$f.json')); w=d.get('resident_window_table') or {}; print('$f', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(w.get('ms_per_step',0),3), round(d['roofline']['frac'],3), d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
