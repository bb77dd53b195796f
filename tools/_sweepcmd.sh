python -m pytest tests -m gpu -x -q -k "window_table or batch or g2_edge or g2_full" 2>&1 | tail -2
python tools/sweep.py --sizes 8,10,12,14,16,17,18,20 --reps 10 --windowed 0 2>&1 | tail -8 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('log2n','ms','c','W','rounds')})"
