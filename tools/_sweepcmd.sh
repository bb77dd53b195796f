for cfg in "lanes=1" "lanes=2" "lanes=3" "lanes=4" "lanes=2 tree_rounds=4" "lanes=2 tree_rounds=5" "lanes=4 tree_rounds=4" "lanes=3 tree_rounds=4"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 14,16,18,19 --reps 10 $args 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms')})"
done
