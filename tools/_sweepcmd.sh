for cfg in "lanes=4 persist=444" "lanes=4 persist=444 ba_k=12" "lanes=4 persist=444 ba_k=16" "lanes=4 persist=444 ba_k=16 pt_k=4" "lanes=4 persist=592 ba_k=16" "lanes=3 persist=444 ba_k=16" "lanes=4 persist=444 ba_k=24"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 18,20 $args 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms')})"
done
