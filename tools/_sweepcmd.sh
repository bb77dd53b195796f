python -m pytest tests -m gpu -x -q -k "window_table" 2>&1 | tail -5
for cfg in "20:16" "20:8" "20:32" "19:16" "20:4"; do
  wb=${cfg%%:*}; ss=${cfg##*:}
  echo "== windowed wb=$wb subslots=$ss"; python tools/sweep.py --sizes 18,20 --reps 8 --windowed $wb --opt subslots=$ss 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','c','W','rounds','adds','k_tree_fwd','k_tree_bwd','k_fold','k_finish','k_sort','k_inv_tree','host_combine')})"
done
