python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "lanes=1" "lanes=2" "lanes=3" "lanes=2 ba_k=16" "lanes=2 ba_k=16 pt_k=8" "lanes=2 ba_k=8 pt_k=8" "lanes=3 ba_k=16 pt_k=8" "lanes=2 ba_k=4 pt_k=4"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 18,20 $args 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','k_tree_fwd','k_inv_tree','k_tree_bwd','k_plan','k_fold','launches')})"
done
