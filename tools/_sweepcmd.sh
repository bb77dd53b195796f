for cfg in "lanes=1" "lanes=2" "lanes=4"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 10,12,13,14,15,16 --reps 20 $args 2>&1 | tail -6 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','rounds')})"
done
