for cv in bls12381 bn128; do
  python tools/sweep.py --curve $cv --sizes 10,12,14,16,18,20,22,24,26 --reps 3 --roofline 2>&1 | grep '^{' 
  python tools/sweep.py --curve $cv --sizes 10,12,14,16,18,20,22,24 --reps 3 --roofline --windowed 0 2>&1 | grep '^{'
done
