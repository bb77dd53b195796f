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
# window table, batch, GLV, G2 and NTT entry points on small inputs
cv = pyref.BLS12_381
bases = make_bases(cv, 3000, 7); sc = make_scalars(3000, 8, "u256")
exp = oracle_msm(cv, bases, sc, 32, 3000)
for wb in (0, 8):
    h = eng.upload_bases_windowed(cv.cid, bases, 3000, 32, wb)
    assert eng.normalize(cv.cid, eng.multiexp_resident(h, sc, 32, 3000, cv.cid)) == exp
    out = eng.multiexp_batch(h, sc * 3, 32, 3000, 3, cv.cid); assert eng.normalize(cv.cid, out[:144]) == exp
    eng.free_bases(h)
p2, s2 = eng.glv_preprocess(cv.cid, bases, sc, 3000)
assert eng.normalize(cv.cid, eng.multiexp_affine(cv.cid, p2, s2, 32, 6000)) == exp
import torch
for cid, n8 in ((2, 96), (3, 64)):
    d = torch.empty(2000 * 2 * n8, dtype=torch.uint8, device="cuda"); eng.generate_bases(cid, 11, 0, 2000, d)
    sd = torch.randint(0, 256, (2000 * 32,), dtype=torch.uint8, device="cuda")
    a = eng.normalize(cid, eng.multiexp_affine(cid, d, sd, 32, 2000))
    h = eng.upload_bases_windowed(cid, d, 2000, 32, 0)
    assert eng.normalize(cid, eng.multiexp_resident(h, sd, 32, 2000, cid)) == a
    eng.free_bases(h)
for cid in (0, 1):
    for lg in (0, 1, 5, 11, 13):
        x = torch.randint(0, 256, (32 << lg,), dtype=torch.uint8, device="cuda"); x.view(-1, 32)[:, 31] &= 0x0F
        y = torch.empty_like(x); eng.fr_fft(cid, x, lg, out=y); eng.fr_fft(cid, y, lg, inverse=True, out=y); eng.synchronize()
        assert torch.equal(x, y), (cid, lg)
# round 2: the grouped-sort path with dense addition items and the cluster fold tail (2^18 points), forced tree rounds on a small input, the warp-tree
# sum of partial points, f1m_batchInverse and the schedule hook
cv = pyref.BLS12_381
n = 1 << 18
d = torch.empty(n * 96, dtype=torch.uint8, device="cuda"); eng.generate_bases(0, 21, 0, n, d)
sd = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
a = eng.normalize(0, eng.multiexp_affine(0, d, sd, 32, n))
eng.set_option("lanes", 1); assert eng.normalize(0, eng.multiexp_affine(0, d, sd, 32, n)) == a; eng.set_option("lanes", 4)
eng.set_option("fold_cluster", 0); assert eng.normalize(0, eng.multiexp_affine(0, d, sd, 32, n)) == a; eng.set_option("fold_cluster", 1)
bases = make_bases(cv, 3000, 7); sc = make_scalars(3000, 8, "u256"); exp = oracle_msm(cv, bases, sc, 32, 3000)
eng.set_option("tree_rounds", 4); eng.set_option("window_bits", 7)
assert eng.normalize(0, eng.multiexp_affine(0, bases, sc, 32, 3000)) == exp
eng.set_option("tree_rounds", -1); eng.set_option("window_bits", 0)
one = eng.multiexp_affine(0, bases, sc, 32, 3000)
for cnt in (1, 2, 8, 33, 70):
    tot = eng.normalize(0, eng.sum_points(0, one * cnt, cnt))
    assert tot == eng.normalize(0, eng.multiexp_affine(0, bases, b"".join(((int.from_bytes(sc[32 * i:32 * i + 32], "little") * cnt) % cv.r).to_bytes(32, "little") for i in range(3000)), 32, 3000)), cnt
xs = b"".join(pyref.fe_bytes(cv, (i * 7919 + 1) % cv.q if i % 11 else 0) for i in range(3000))
assert eng.fq_batch_inverse(0, xs) == eng.fq_op(0, 4, xs)
plan, offs, srt = eng.debug_schedule(sc[:32 * 500], 32, 500, 9); assert offs[-1] == len(srt)
print("sanitizer case ok")
