"""tiny end-to-end case for compute-sanitizer: both curves, batch-affine + serial accumulate, codecs, chunk API"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zprize-wasm-msm_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import b200msm, pyref, coracle
from util import make_bases, make_scalars, oracle_msm
eng = b200msm.Engine(0)
for cname in ("bls12381", "bn128"):
    cv = pyref.CURVES[cname]
    for n in (1, 37, 3000, 70000):
        bases = make_bases(cv, n, 5) if n <= 3000 else make_bases(cv, 3000, 5) * 24
        m = len(bases) // (2 * cv.n8); sc = make_scalars(m, n, "u256")
        got = eng.normalize(cv.cid, eng.multiexp_affine(cv.cid, bases, sc, 32, m))
        assert got == oracle_msm(cv, bases, sc, 32, m), (cname, n)
    eng.set_option("accumulate", 1)
    bases = make_bases(cv, 500, 6); sc = make_scalars(500, 3, "u256")
    assert eng.normalize(cv.cid, eng.multiexp_affine(cv.cid, bases, sc, 32, 500)) == oracle_msm(cv, bases, sc, 32, 500)
    eng.set_option("accumulate", 0)
    ch = eng.multiexp_affine_chunk(cv.cid, bases, sc, 32, 500, 250, 11)
    assert eng.normalize(cv.cid, ch) == coracle.normalize(cv.cid, coracle.multiexp_affine_chunk(cv.cid, bases, sc, 32, 500, 250, 11))
    c = eng.batch_convert(cv.cid, "LEMtoC", bases, 500); assert eng.batch_convert(cv.cid, "CtoLEM", c, 500) == bases
print("sanitizer case ok")
