python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for cfg in "tree_rounds=-1" "tree_rounds=2" "tree_rounds=3" "tree_rounds=4"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 14,16,18,20 $args 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','k_finish')})"
done
