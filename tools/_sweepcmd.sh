for cfg in "lanes=4" "lanes=6" "lanes=8" "lanes=8 persist=296" "lanes=6 persist=592" "lanes=4 persist=592" "lanes=4 persist=296"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 16,18,20,22 --reps 8 $args 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms')})"
done
