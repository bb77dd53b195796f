for cfg in "accumulate=1 window_bits=10" "accumulate=1 window_bits=12" "accumulate=1 window_bits=13" "accumulate=1 window_bits=14" "accumulate=1 window_bits=15" "accumulate=2"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"; python tools/sweep.py --sizes 12,14,16,18 --reps 8 $args 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','c','W','k_sort','k_fold','accumulate','bucket_reduce','window_combine')})"
done
