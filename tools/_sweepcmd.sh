python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cv in bls12381_g2 bn128_g2; do
for cfg in "-1" "0"; do
  echo "== $cv windowed $cfg"; python tools/sweep.py --curve $cv --sizes 16,18,20 --reps 5 --windowed $cfg 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','c','W','rounds','adds','k_tree_fwd','k_tree_bwd','k_fold','k_finish','k_sort','k_inv_tree','host_combine')})"
done; done
python tools/sweep.py --curve bls12381 --sizes 18,20 --reps 8 2>&1 | tail -2 | cut -c1-200
