"""CPU tests: pin the oracles (oracle/pyref.py, oracle/msm_oracle.c) against the reference's golden
vectors (tests/golden/*.json, from wasmcurves/test/*.js) and against the reference's own WASM module
compiled natively (oracle/_ref, when built)."""
import json, os, random
import pytest
import pyref, coracle, refwasm

G = os.path.join(os.path.dirname(__file__), "golden")
BA = json.load(open(os.path.join(G, "batchAffine.json")))["tests"]
GLV = json.load(open(os.path.join(G, "glv.json")))["tests"]
BLS = pyref.BLS12_381


def ints(v):
    return [ints(x) for x in v] if isinstance(v, list) else int(v, 16)


def vals(name, d=BA):
    return {k: ints(v) for k, v in d[name]["values"].items()}


def pts_of(flat):
    return [(flat[2 * i], flat[2 * i + 1]) if (flat[2 * i] or flat[2 * i + 1]) else None for i in range(len(flat) // 2)]


def sched_sum(cv, points, sched):
    acc = None
    for w in sched:
        if w == pyref.SENTINEL: continue
        acc = pyref.add(cv, acc, pyref.mul(cv, w & 0x7FFFFFFF, points[w >> 32]))
    return acc


# ---- schedule KATs (test/batchAffine.js:43-258)
def test_get_chunk_kat():
    v = vals("getChunk is correct.")
    for s, c, sz, nch, exp in zip(v["inputScalarArr"], v["chunkSizeArr"], v["scalarSizeArr"], v["numChunksArr"], v["expectedOutputArr"]):
        got = [pyref.get_chunk(s.to_bytes(8, "little")[:sz], sz, k * c, c) for k in range(nch)]
        assert got == exp


def test_compute_schedule_kat():
    v = vals("computeSchedule is correct.")
    sb = b"".join(s.to_bytes(v["scalarSize"], "little") for s in v["inputScalars"])
    sched, counts = pyref.compute_schedule(sb, v["numPoints"], v["chunkSize"], v["scalarSize"])
    assert sched == [w for row in v["expectedOutputPointSchedules"] for w in row]
    assert counts == v["expectedOutputRoundCounts"]


def test_organize_buckets_kat():
    v = vals("organizeBuckets is correct.")
    n, W, B = v["numPoints"], v["numChunks"], v["numBuckets"]
    out, cnts = [], []
    c = B.bit_length() - 1
    for k in range(W):
        o, cn = pyref.organize_buckets_one_round(v["inputs"][k * n:(k + 1) * n], c)
        out += o; cnts += cn
    assert out == v["expectedOutput"] and cnts == v["expectedOutputBucketCounts"]


# ---- group-level KATs
def test_points_are_multiples_of_generator():
    v = vals("multiExp is correct (case 1).")
    P = pts_of(v["inputPoints"])
    for k, p in zip([1, 2, 3, 5, 4, 6, 7, 8, 9, 9], P):
        assert p == pyref.mul(BLS, k, BLS.G)
        assert pyref.is_on_curve(BLS, p)


def test_multiexp_case1_kat_all_oracles():
    """test/batchAffine.js:1177-1255 -- the reference's one end-to-end MSM KAT."""
    v = vals("multiExp is correct (case 1).")
    P = pts_of(v["inputPoints"]); n = v["numPoints"]
    sb = b"".join(s.to_bytes(32, "little") for s in v["inputScalars"])
    exp = tuple(v["expectedOutput"])
    assert pyref.msm_naive(BLS, P, v["inputScalars"]) == exp
    assert pyref.multiexp_affine(BLS, P, sb, 32, n) == exp
    bases = b"".join(pyref.affine_to_bytes(BLS, p) for p in P)
    assert coracle.normalize(0, coracle.multiexp_affine(0, bases, sb, 32, n)) == pyref.canonical_bytes(BLS, exp)
    if refwasm.available("bls12381"):
        assert refwasm.RefModule("bls12381").msm_affine(bases, sb, 32, n) == exp


def test_single_chunk_kat():
    v = vals("multiExpSingleChunk is correct.")
    assert sched_sum(BLS, pts_of(v["inputPoints"]), v["pointSchedules"]) == tuple(v["expectedOutput"])


@pytest.mark.parametrize("case", ["multiExpChunks is correct (case 1).", "multiExpChunks is correct (case 2)."])
def test_multiexp_chunks_kat(case):
    v = vals(case)
    P = pts_of(v["inputPoints"]); c = v["chunkSize"]
    acc = None
    for k in reversed(range(v["numChunks"])):      # top window first, Horner (build_multiexp_opt.js:1823-1954)
        for _ in range(c): acc = pyref.add(BLS, acc, acc)
        acc = pyref.add(BLS, acc, sched_sum(BLS, P, v["pointSchedules"][k]))
    assert acc == tuple(v["expectedOutput"])


def test_accumulate_across_chunks_kat():
    v = vals("accumulateAcrossChunks is correct.")
    A = pts_of(v["inputAccumulator"]); S = pts_of(v["inputAccumulatorSingleChunk"]); E = pts_of(v["expectedOutput"])
    # first test: top chunk, no doubling; second: 2^5 * acc + chunk
    assert pyref.add(BLS, A[0], S[0]) == E[0]
    assert pyref.add(BLS, pyref.mul(BLS, 1 << v["chunkSize"], A[1]), S[1]) == E[1]


def test_reduce_buckets_kat():
    v = vals("reduceBuckets is correct.")
    P = pts_of(v["inputPoints"]); E = pts_of(v["expectedOutput"])
    by_bucket = {}
    for w in v["pointSchedules"]:
        if w == pyref.SENTINEL: continue
        by_bucket.setdefault(w & 0x7FFFFFFF, []).append(P[w >> 32])
    sums = []
    for b in sorted(by_bucket):
        acc = None
        for p in by_bucket[b]: acc = pyref.add(BLS, acc, p)
        sums.append(acc)
    assert sums == E[:len(sums)]


def test_glv_decompose_kat():
    """test/glv.js:50-65: exact |k1|, |k2| and sign of g1m_glv_decomposeScalar; k = k1 - k2*lambda (mod r)."""
    v = vals("decomposeScalar is correct.", GLV)
    k = v["scalar"]
    k1, k2, sign = pyref.glv_decompose(k)
    assert [k1, k2] == v["expectedOutput"] and sign == 1
    s1 = k1 if sign & 1 else -k1; s2 = k2 if sign & 2 else -k2
    assert (s1 + s2 * pyref.GLV_LAMBDA) % BLS.r == k % BLS.r


def test_glv_preprocess_kat():
    """test/glv.js:103-192: scalar halves of g1m_glv_preprocessEndomorphism, and the MSM over the 2N outputs equals the
    MSM over the N inputs."""
    v = vals("preprocessEndomorphism is correct.", GLV)
    n = v["numPoints"]; P = pts_of(v["inputPoints"])[:n]; S = v["inputScalars"][:n]
    P2, S2 = pyref.glv_preprocess(P, S)
    assert S2 == v["expectedScalarOutput"][:2 * n]
    assert all(pyref.is_on_curve(BLS, p) for p in P2)
    assert pyref.msm_naive(BLS, P2, S2) == pyref.msm_naive(BLS, P, S)


def test_glv_decompose_edge_scalars():
    """the decomposition holds (mod r) for every 256-bit scalar, including those >= r, and both halves fit in 128 bits"""
    rnd = random.Random(77)
    edge = [0, 1, 2, BLS.r - 1, BLS.r, BLS.r + 1, 2 * BLS.r - 1, 2 * BLS.r, 2 * BLS.r + 1, (1 << 256) - 1, 1 << 128, (1 << 128) - 1, pyref.GLV_LAMBDA]
    for k in edge + [rnd.getrandbits(256) for _ in range(300)]:
        k1, k2, sign = pyref.glv_decompose(k)
        assert k1 < (1 << 128) and k2 < (1 << 128)
        s1 = k1 if sign & 1 else -k1; s2 = k2 if sign & 2 else -k2
        assert (s1 + s2 * pyref.GLV_LAMBDA) % BLS.r == k % BLS.r


# ---- oracle <-> reference WASM differential (random inputs, both curves)
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("n", [1, 3, 100, 1000])
def test_oracles_agree_random(cname, n):
    cv = pyref.CURVES[cname]
    gen = pyref.affine_to_bytes(cv, cv.G)
    bases = coracle.generate_bases(cv.cid, gen, 0xB2000000 + n, 0, n)
    rnd = random.Random(n)
    sc = [rnd.randrange(0, 1 << 256) for _ in range(n)]
    sb = b"".join(s.to_bytes(32, "little") for s in sc)
    c_res = coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, bases, sb, 32, n))
    if n <= 100:
        P = [pyref.affine_from_bytes(cv, bases[i * 2 * cv.n8:(i + 1) * 2 * cv.n8]) for i in range(n)]
        assert all(pyref.is_on_curve(cv, p) for p in P)
        assert c_res == pyref.canonical_bytes(cv, pyref.msm_naive(cv, P, sc))
    if refwasm.available(cname):
        pb = refwasm.RefModule(cname)
        assert c_res == pyref.canonical_bytes(cv, pb.msm_affine(bases, sb, 32, n))
        # per-window export (_chunk), incl. the truncated top window
        for start, bits in ((0, 5), (13, 7), (250, 11)):
            a = coracle.normalize(cv.cid, coracle.multiexp_affine_chunk(cv.cid, bases, sb, 32, n, start, bits))
            assert a == pyref.canonical_bytes(cv, pb.msm_chunk(bases, sb, 32, n, start, bits))


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_bn128_style_sum_i_times_iG(cname):
    """test/bn128.js:462-497: sum_{i=1..8} i*(i*G) == 204*G."""
    cv = pyref.CURVES[cname]
    P = [pyref.mul(cv, i, cv.G) for i in range(1, 9)]
    sb = b"".join(i.to_bytes(32, "little") for i in range(1, 9))
    bases = b"".join(pyref.affine_to_bytes(cv, p) for p in P)
    exp = pyref.canonical_bytes(cv, pyref.mul(cv, 204, cv.G))
    assert coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, bases, sb, 32, 8)) == exp
    if refwasm.available(cname):
        assert pyref.canonical_bytes(cv, refwasm.RefModule(cname).msm_affine(bases, sb, 32, 8)) == exp


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_group_order_and_edge_cases(cname):
    """test/bls12381.js:339-347 (r*G = 0) + n = 0, zero scalars, infinity inputs, P + (-P)."""
    cv = pyref.CURVES[cname]
    assert pyref.mul(cv, cv.r, cv.G) is None
    g = pyref.affine_to_bytes(cv, cv.G)
    zero = bytes(2 * cv.n8)
    assert coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, b"", b"", 32, 0)) == zero
    assert coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, g * 4, bytes(32 * 4), 32, 4)) == zero
    # r*G through the MSM path
    assert coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, g, cv.r.to_bytes(32, "little"), 32, 1)) == zero
    # G*5 + (-G)*5 = 0 ; infinity input ignored
    ng = pyref.affine_to_bytes(cv, pyref.neg(cv, cv.G))
    five = (5).to_bytes(32, "little")
    assert coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, g + ng + zero, five * 3, 32, 3)) == zero


def test_field_ops_against_python():
    rnd = random.Random(7)
    for cv in (pyref.BLS12_381, pyref.BN254):
        edge = [0, 1, 2, cv.q - 1, cv.q - 2, (cv.q - 1) // 2, cv.R % cv.q]       # test/f1.js:294-342 style edge values
        xs = edge + [rnd.randrange(cv.q) for _ in range(50)]
        ys = list(reversed(edge)) + [rnd.randrange(cv.q) for _ in range(50)]
        a = b"".join(pyref.fe_bytes(cv, x) for x in xs); b = b"".join(pyref.fe_bytes(cv, y) for y in ys)
        Ri = pow(cv.R, -1, cv.q)
        assert coracle.fe_mul(cv.cid, a, b) == b"".join(pyref.fe_bytes(cv, x * y * Ri % cv.q) for x, y in zip(xs, ys))
        assert coracle.fe_add(cv.cid, a, b) == b"".join(pyref.fe_bytes(cv, (x + y) % cv.q) for x, y in zip(xs, ys))
        assert coracle.fe_sub(cv.cid, a, b) == b"".join(pyref.fe_bytes(cv, (x - y) % cv.q) for x, y in zip(xs, ys))
        assert coracle.fe_to_mont(cv.cid, a) == b"".join(pyref.fe_bytes(cv, pyref.to_mont(cv, x)) for x in xs)
        inv = coracle.fe_inv(cv.cid, a)
        for i, x in enumerate(xs):  # Montgomery-domain inverse: (xR)^-1 * R^2 ... checked via mul == one
            got = pyref.fe_from(cv, inv[i * cv.n8:(i + 1) * cv.n8])
            assert (got == 0) if x == 0 else (got * x * Ri % cv.q == cv.R % cv.q)


# ---- G2: the Python restatement over Fq2 against the reference's own g2m exports
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_g2_restatement_matches_reference_wasm(cname):
    """g2m_multiexpAffine / g2m_timesScalarAffine of the reference module vs affine arithmetic over Fq2 = Fq[u]/(u^2+1)"""
    if not refwasm.available(cname): pytest.skip("oracle/_ref not built")
    cv = pyref.CURVES[cname]
    g2 = refwasm.RefG2(refwasm.RefModule(cname))
    G = pyref.g2_from_bytes(cv, g2.generator_affine())
    rnd = random.Random(31)
    ks = [rnd.getrandbits(64) | 1 for _ in range(6)]
    bases = [g2.times_scalar_affine(g2.generator_affine(), k.to_bytes(8, "little")) for k in ks]
    P = [pyref.g2_from_bytes(cv, b) for b in bases]
    assert P == [pyref.g2_mul(cv, k, G) for k in ks]
    sc = [rnd.getrandbits(256) for _ in ks]
    got = g2.msm_affine(b"".join(bases), b"".join(s.to_bytes(32, "little") for s in sc), 32, len(ks))
    assert got == pyref.g2_canonical_bytes(cv, pyref.g2_msm_naive(cv, P, sc))
    assert pyref.g2_mul(cv, cv.r, G) is None                                   # test/bls12381.js:349-357 style: r * G2 = 0
