#!/usr/bin/env python3
"""Instruction mix of the hot kernels, read from the SASS of the shipped library (no GPU needed):

    python tools/sass_counts.py [--lib zprize-wasm-msm_b200/b200msm/libb200msm.so] [--out profiles/r2_sass_counts.txt]

For every kernel whose name matches one of --kernels: registers / stack / shared memory (cuobjdump -res-usage) and the
counts of IMAD.WIDE (the 32x32+64 multiply-add the field multiplier is made of), other IMAD forms, IADD3, LDG / STG and
the total -- the numbers DESIGN.md section 6 quotes for the instruction bound of k_tree_bwd (5 field multiplications x
300 limb products = 1500 IMAD.WIDE per addition).  Also a probe kernel that is one field multiplication per loop trip.
"""
import argparse, collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "zprize-wasm-msm_b200", "b200msm", "libb200msm.so"))
ap.add_argument("--kernels", default="k_tree_bwd,k_tree_fwd,k_tree_meta,k_fold,k_accum_finish,k_digits,k_inv_root,k_prod_fwd,k_prod_bwd,k_fpmul_probe,k_ntt")
ap.add_argument("--curves", default="BLS12_381,BN254", help="substrings of the mangled template arguments to keep (G2 = Fq2 instantiations are skipped unless named)")
ap.add_argument("--out", default="")
a = ap.parse_args()
cuobjdump = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


res = subprocess.run([cuobjdump, "-res-usage", a.lib], capture_output=True, text=True, check=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
    usage[m.group(1)] = m.group(2).strip()
want = [k for k in a.kernels.split(",") if k]
curves = [c for c in a.curves.split(",") if c]
funcs = [f for f in usage if any(k in f for k in want) and (not curves or any(c in f for c in curves) or "IN" not in f) and ("Fq2" in a.curves or "3Fq2" not in f)]
pretty = demangle(funcs)
lines = ["# SASS instruction mix of the hot kernels -- %s" % os.path.relpath(a.lib, ROOT),
         "# produced by tools/sass_counts.py (cuobjdump -sass -fun <kernel>); IMAD.WIDE = 32x32+64 multiply-add (half rate), IMAD* = every IMAD form incl. WIDE",
         "# %-92s %5s %5s %6s %6s %6s %5s %5s %6s  %s" % ("kernel", "WIDE", "IMAD*", "IADD3", "LDG", "STG", "LDL", "STL", "total", "resources")]
for f in sorted(funcs, key=lambda x: pretty[x]):
    sass = subprocess.run([cuobjdump, "-sass", "-fun", f, a.lib], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m: ops[m.group(1)] += 1
    tot = sum(ops.values())
    wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
    imad = sum(v for k, v in ops.items() if k.startswith("IMAD"))
    g = lambda p: sum(v for k, v in ops.items() if k.startswith(p))
    name = re.sub(r"^void |b200::|\(anonymous namespace\)::", "", pretty[f]).split("(")[0]
    lines.append("%-94s %5d %5d %6d %6d %6d %5d %5d %6d  %s" % (name[:94], wide, imad, g("IADD3"), g("LDG"), g("STG"), g("LDL"), g("STL"), tot, " ".join(x for x in usage[f].split() if x.startswith(("REG", "STACK", "SHARED")))))
text = "\n".join(lines) + "\n"
if a.out:
    open(os.path.join(ROOT, a.out) if not os.path.isabs(a.out) else a.out, "w").write(text)
sys.stdout.write(text)
