for cfg in "ba_k=8" "ba_k=12" "ba_k=16" "ba_k=16 pt_k=4" "ba_k=8 pt_k=16"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 18,20,22 --reps 8 $args 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms')})"
done
