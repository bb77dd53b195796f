python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "-1" "0"; do
  echo "== windowed $cfg"; python tools/sweep.py --sizes 14,16,18,20,22 --reps 8 --windowed $cfg 2>&1 | tail -5 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','c','W','rounds','adds','k_tree_fwd','k_tree_bwd','k_fold','k_finish','k_sort','k_inv_tree')})"
done
