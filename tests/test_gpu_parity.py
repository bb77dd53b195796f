"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the oracles
on the same seeded inputs -- bit-exact on canonical affine output (SURVEY.md 8c parity definition)."""
import json, os, random
import pytest
import pyref, coracle, refwasm
from util import curve, make_bases, make_scalars, oracle_msm, gen_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import b200msm
    e = b200msm.Engine()
    yield e
    e.close()


def msm(eng, cv, bases, scalars, ssz, n):
    return eng.normalize(cv.cid, eng.multiexp_affine(cv.cid, bases, scalars, ssz, n))


# ---------------------------------------------------------------- field kernels (build_f1m.js)
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_fq_ops_bit_exact(eng, cname):
    cv = curve(cname); rnd = random.Random(11)
    edge = [0, 1, 2, cv.q - 1, cv.q - 2, (cv.q - 1) // 2, cv.R % cv.q, (cv.R * cv.R) % cv.q, (1 << (8 * cv.n8 - 3)) % cv.q]
    xs = edge + [rnd.randrange(cv.q) for _ in range(3000)]
    ys = list(reversed(edge)) + [rnd.randrange(cv.q) for _ in range(3000)]
    a = b"".join(pyref.fe_bytes(cv, x) for x in xs); b = b"".join(pyref.fe_bytes(cv, y) for y in ys)
    assert eng.fq_op(cv.cid, 0, a, b) == coracle.fe_mul(cv.cid, a, b)
    assert eng.fq_op(cv.cid, 1, a, b) == coracle.fe_add(cv.cid, a, b)
    assert eng.fq_op(cv.cid, 2, a, b) == coracle.fe_sub(cv.cid, a, b)
    assert eng.fq_op(cv.cid, 3, a) == coracle.fe_mul(cv.cid, a, a)
    assert eng.fq_op(cv.cid, 5, a) == coracle.fe_to_mont(cv.cid, a)
    assert eng.fq_op(cv.cid, 6, a) == coracle.fe_from_mont(cv.cid, a)
    assert eng.fq_op(cv.cid, 7, a) == coracle.fe_sub(cv.cid, bytes(len(a)), a)
    small = a[:200 * cv.n8]
    inv = coracle.fe_inv(cv.cid, small)
    assert eng.fq_op(cv.cid, 4, small) == inv          # binary extended Euclid
    assert eng.fq_op(cv.cid, 8, small) == inv          # Fermat


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_radix29_multiplier_bit_exact(eng, cname):
    """fp29.cuh: carry-free radix-2^29 Montgomery product (R' = 2^(29L)) and its change of radix to/from the reference's R.
    A rejected round-1 alternative: compiled only into -DB200_EXPERIMENTS builds (B200_EXPERIMENTS=1 python __graft_entry__.py)."""
    import b200msm
    if not hasattr(b200msm.lib, "b200msm_probe_dfma"): pytest.skip("not an experiments build")
    cv = curve(cname); rnd = random.Random(29)
    L = (cv.q.bit_length() + 28) // 29; Rp = 1 << (29 * L)
    edge = [0, 1, 2, cv.q - 1, cv.q - 2, (cv.q - 1) // 2, cv.R % cv.q, Rp % cv.q]
    xs = edge + [rnd.randrange(cv.q) for _ in range(3000)]
    ys = list(reversed(edge)) + [rnd.randrange(cv.q) for _ in range(3000)]
    a = b"".join(pyref.fe_bytes(cv, x) for x in xs); b = b"".join(pyref.fe_bytes(cv, y) for y in ys)
    Rpi = pow(Rp, -1, cv.q)
    assert eng.fq_op(cv.cid, 9, a, b) == b"".join(pyref.fe_bytes(cv, x * y * Rpi % cv.q) for x, y in zip(xs, ys))
    assert eng.fq_op(cv.cid, 10, a, b) == coracle.fe_mul(cv.cid, a, b)      # enter -> mul -> leave == f1m_mul


# ---------------------------------------------------------------- synthetic bases generator
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_generate_bases_matches_oracle(eng, cname):
    import torch
    cv = curve(cname); n = 300
    d = torch.empty(n * 2 * cv.n8, dtype=torch.uint8, device="cuda")
    eng.generate_bases(cv.cid, 0xB2000000, 5, n, d)
    got = bytes(d.cpu().numpy())
    exp = coracle.generate_bases(cv.cid, gen_bytes(cv), 0xB2000000, 5, n)
    assert got == exp


# ---------------------------------------------------------------- MSM vs oracle, both accumulate forms
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n", [1, 2, 3, 10, 100, 1000, 5000])
def test_msm_random_matches_oracle(eng, cname, mode, n):
    cv = curve(cname)
    eng.set_option("accumulate", mode)
    try:
        bases = make_bases(cv, n); sc = make_scalars(n, 1000 + n, "u256")
        assert msm(eng, cv, bases, sc, 32, n) == oracle_msm(cv, bases, sc, 32, n)
    finally:
        eng.set_option("accumulate", 0)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("wb", [1, 2, 5, 9, 13, 16])
def test_msm_window_widths(eng, cname, wb):
    cv = curve(cname); n = 700
    eng.set_option("window_bits", wb)
    try:
        bases = make_bases(cv, n, 77); sc = make_scalars(n, 5, "u256")
        assert msm(eng, cv, bases, sc, 32, n) == oracle_msm(cv, bases, sc, 32, n)
    finally:
        eng.set_option("window_bits", 0)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("wb", [1, 7, 16])
def test_device_window_combination_matches_host_tail(eng, cname, wb):
    """combine=1 runs k_window_sums + k_horner on the device instead of the host-side serial tail: same point."""
    cv = curve(cname); n = 900
    bases = make_bases(cv, n, 41); sc = make_scalars(n, 42, "u256")
    exp = oracle_msm(cv, bases, sc, 32, n)
    eng.set_option("window_bits", wb)
    try:
        for mode in (1, 0):
            eng.set_option("combine", mode)
            assert msm(eng, cv, bases, sc, 32, n) == exp, mode
    finally:
        eng.set_option("combine", 0); eng.set_option("window_bits", 0)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("lanes,group_pairs", [(1, 0), (2, 0), (4, 0), (3, 9000), (1, 30000), (4, 2000)])
def test_window_groups_and_lanes(eng, cname, lanes, group_pairs):
    """Window slots are processed in groups (memory budget) that alternate between stream lanes; any grouping gives the same point."""
    cv = curve(cname); n = 6000
    bases = make_bases(cv, n, 51); sc = make_scalars(n, 52, "u256")
    exp = oracle_msm(cv, bases, sc, 32, n)
    eng.set_option("lanes", lanes); eng.set_option("group_pairs", group_pairs)
    try:
        assert msm(eng, cv, bases, sc, 32, n) == exp
        # skewed input: only the two lowest windows are populated, the top groups are empty
        sc2 = b"".join((int.from_bytes(sc[i * 32:i * 32 + 3], "little")).to_bytes(32, "little") for i in range(n))
        assert msm(eng, cv, bases, sc2, 32, n) == oracle_msm(cv, bases, sc2, 32, n)
    finally:
        eng.set_option("lanes", 4); eng.set_option("group_pairs", 0)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_msm_edge_cases(eng, cname):
    cv = curve(cname); n8 = cv.n8
    zero = bytes(2 * n8)
    # n = 0 -> canonical zero (build_multiexp.js:283-287)
    assert msm(eng, cv, b"", b"", 32, 0) == zero
    raw0 = eng.multiexp_affine(cv.cid, b"", b"", 32, 0)
    assert raw0 == bytes(n8) + pyref.fe_bytes(cv, cv.R % cv.q) + bytes(n8)
    n = 600
    bases = make_bases(cv, n, 9)
    for kind in ("small", "equal"):
        sc = make_scalars(n, 3, kind)
        assert msm(eng, cv, bases, sc, 32, n) == oracle_msm(cv, bases, sc, 32, n), kind
    assert msm(eng, cv, bases, bytes(32 * n), 32, n) == zero                      # all-zero scalars
    # scalars >= r, r - 1, 2^256 - 1 (scalars are plain integers, not reduced)
    special = [cv.r - 1, cv.r, cv.r + 1, (1 << 256) - 1, 1 << 255, 1, 0, (1 << 128) - 1]
    sc = b"".join(v.to_bytes(32, "little") for v in special)
    b8 = bases[:len(special) * 2 * n8]
    assert msm(eng, cv, b8, sc, 32, len(special)) == oracle_msm(cv, b8, sc, 32, len(special))
    # duplicate points, P and -P, infinity inputs: exercises doubling / cancellation / skip paths of batch-affine
    g = gen_bytes(cv); ng = pyref.affine_to_bytes(cv, pyref.neg(cv, cv.G))
    pts = (g + g + ng + zero + g + ng + ng + zero) * 40
    m = len(pts) // (2 * n8)
    for seed, kind in ((1, "equal"), (2, "small"), (3, "u256")):
        sc = make_scalars(m, seed, kind)
        assert msm(eng, cv, pts, sc, 32, m) == oracle_msm(cv, pts, sc, 32, m), kind
    # G*5 + (-G)*5 = 0
    five = (5).to_bytes(32, "little")
    assert msm(eng, cv, g + ng, five * 2, 32, 2) == zero


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("ssz", [1, 4, 5, 16, 31])
def test_msm_scalar_sizes(eng, cname, ssz):
    cv = curve(cname); n = 300
    bases = make_bases(cv, n, 21); rnd = random.Random(ssz)
    sc = bytes(rnd.getrandbits(8) for _ in range(n * ssz))
    assert msm(eng, cv, bases, sc, ssz, n) == oracle_msm(cv, bases, sc, ssz, n)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_chunk_api_matches_oracle(eng, cname):
    """g1m_multiexpAffine_chunk (build_multiexp.js:96-249), incl. the clipped top window and Horner recombination."""
    cv = curve(cname); n = 500
    bases = make_bases(cv, n, 33); sc = make_scalars(n, 8, "u256")
    for start, bits in ((0, 5), (13, 7), (250, 11), (100, 16), (255, 1), (248, 8)):
        got = eng.normalize(cv.cid, eng.multiexp_affine_chunk(cv.cid, bases, sc, 32, n, start, bits))
        exp = coracle.normalize(cv.cid, coracle.multiexp_affine_chunk(cv.cid, bases, sc, 32, n, start, bits))
        assert got == exp, (start, bits)
    # Horner over chunks == whole MSM (the caller-side combination, build_multiexp.js:319-369)
    c = 13; acc = None
    nch = (256 - 1) // c + 1
    for k in reversed(range(nch)):
        if acc is not None:
            for _ in range(c): acc = pyref.add(cv, acc, acc)
        ch = eng.normalize(cv.cid, eng.multiexp_affine_chunk(cv.cid, bases, sc, 32, n, k * c, c))
        x = int.from_bytes(ch[:cv.n8], "little"); y = int.from_bytes(ch[cv.n8:], "little")
        acc = pyref.add(cv, acc, None if (x == 0 and y == 0) else (x, y))
    assert pyref.canonical_bytes(cv, acc) == oracle_msm(cv, bases, sc, 32, n)


def test_reference_kat_through_protoboard(eng):
    """The reference's end-to-end KAT (test/batchAffine.js:1177-1255), written the way the reference writes it."""
    import b200msm
    G = os.path.join(os.path.dirname(__file__), "golden")
    v = json.load(open(os.path.join(G, "batchAffine.json")))["tests"]["multiExp is correct (case 1)."]["values"]
    inputPoints = [int(x, 16) for x in v["inputPoints"]]; inputScalars = [int(x, 16) for x in v["inputScalars"]]
    expectedOutput = [int(x, 16) for x in v["expectedOutput"]]
    numPoints = 10; n8q = 48; n8r = 32
    pb = b200msm.Protoboard("bls12381", engine=eng)
    pRes = pb.alloc(n8q * 3); pPoints = pb.alloc(numPoints * n8q * 2); pScalars = pb.alloc(numPoints * n8r)
    for i in range(numPoints):
        pb.set(pPoints + 96 * i, inputPoints[i * 2], 48); pb.set(pPoints + 96 * i + 48, inputPoints[i * 2 + 1], 48)
        pb.f1m_toMontgomery(pPoints + 96 * i, pPoints + 96 * i); pb.f1m_toMontgomery(pPoints + 96 * i + 48, pPoints + 96 * i + 48)
    for i in range(numPoints):
        pb.set(pScalars + n8r * i, inputScalars[i], n8r)
    pb.g1m_multiexp_multiExp(pPoints, pScalars, numPoints, pRes)
    pb.g1m_normalize(pRes, pRes)
    pb.f1m_fromMontgomery(pRes, pRes); pb.f1m_fromMontgomery(pRes + 48, pRes + 48)
    output = pb.get(pRes, 2, 48)
    assert output[0] == expectedOutput[0] and output[1] == expectedOutput[1]


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_bn128_style_multiexp(eng, cname):
    """test/bn128.js:462-497: sum_{i=1..8} i * (i*G) = 204*G"""
    cv = curve(cname)
    P = [pyref.mul(cv, i, cv.G) for i in range(1, 9)]
    bases = b"".join(pyref.affine_to_bytes(cv, p) for p in P); sc = b"".join(i.to_bytes(32, "little") for i in range(1, 9))
    assert msm(eng, cv, bases, sc, 32, 8) == pyref.canonical_bytes(cv, pyref.mul(cv, 204, cv.G))


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_msm_2_14_against_reference_wasm(eng, cname):
    """BASELINE configs[0]: 2^14 points; checked against the reference's own WASM module run natively."""
    cv = curve(cname); n = 1 << 14
    bases = make_bases(cv, n, 0xB2000000 + 14); sc = make_scalars(n, 14, "u256")
    got = msm(eng, cv, bases, sc, 32, n)
    assert got == oracle_msm(cv, bases, sc, 32, n)
    if refwasm.available(cname):
        assert got == pyref.canonical_bytes(cv, refwasm.RefModule(cname).msm_affine(bases, sc, 32, n))


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_sum_and_resident_api(eng, cname):
    cv = curve(cname); n = 2000
    bases = make_bases(cv, n, 5); sc = make_scalars(n, 6, "modr", cv.r)
    h = eng.upload_bases(cv.cid, bases, n)
    try:
        full, st = eng.multiexp_resident(h, sc, 32, n, cv.cid, want_stats=True)
        assert eng.normalize(cv.cid, full) == oracle_msm(cv, bases, sc, 32, n)
        assert st["n"] == n and st["pairs"] > 0 and st["ms_total"] > 0
        # point-range shards + g1m_add merge
        half = n // 2
        a = eng.multiexp_affine(cv.cid, bases[:half * 2 * cv.n8], sc[:half * 32], 32, half)
        b = eng.multiexp_affine(cv.cid, bases[half * 2 * cv.n8:], sc[half * 32:], 32, n - half)
        assert eng.normalize(cv.cid, eng.sum_points(cv.cid, a + b, 2)) == eng.normalize(cv.cid, full)
    finally:
        eng.free_bases(h)


# ---------------------------------------------------------------- BASELINE.json full sizes: exact known answers
def _splitmix64_np(x):
    import numpy as np
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


@pytest.mark.parametrize("cname,lg", [("bls12381", 16), ("bls12381", 18), ("bls12381", 20), ("bn128", 20)])
def test_full_size_known_answer(eng, cname, lg):
    """BASELINE configs at full size.  The synthetic bases are P_i = k_i*G with known k_i (splitmix64 stream), so the
    exact answer of sum_i s_i*P_i is (sum_i s_i*k_i mod r)*G -- computed with host big integers and ONE oracle scalar
    multiplication, independent of any MSM code.  Bit-exact on canonical affine output, plus linearity and sharding."""
    import numpy as np, torch
    cv = curve(cname); n = 1 << lg; n8 = cv.n8
    seed = 0xB2000000 + lg
    d = torch.empty(n * 2 * n8, dtype=torch.uint8, device="cuda")
    eng.generate_bases(cv.cid, seed, 0, n, d)
    with np.errstate(over="ignore"):
        k = _splitmix64_np(np.uint64(seed) + np.arange(n, dtype=np.uint64))
    k[k == 0] = 1
    rng = np.random.default_rng(lg)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x7F                                  # < 2^255 so that s1 + s2 below stays a 256-bit integer
    sb = sc.tobytes()
    sd = torch.from_numpy(sc.reshape(-1).copy()).cuda()
    got = eng.normalize(cv.cid, eng.multiexp_affine(cv.cid, d, sd, 32, n))
    # exact expectation
    words = sc.view("<u8").astype(object)              # n x 4 little-endian 64-bit words
    svals = words[:, 0] + (words[:, 1] << 64) + (words[:, 2] << 128) + (words[:, 3] << 192)
    total = int((svals * k.astype(object)).sum() % cv.r)
    exp = coracle.normalize(cv.cid, coracle.times_scalar_affine(cv.cid, gen_bytes(cv), total.to_bytes(32, "little")))
    assert got == exp
    # sharding property at full size: two halves summed == whole
    h = n // 2
    a = eng.multiexp_affine(cv.cid, d[: h * 2 * n8], sd[: h * 32], 32, h)
    b = eng.multiexp_affine(cv.cid, d[h * 2 * n8:], sd[h * 32:], 32, n - h)
    assert eng.normalize(cv.cid, eng.sum_points(cv.cid, a + b, 2)) == exp
    # linearity: MSM(s) + MSM(s') == MSM(s + s') for a second scalar vector (s' = byte-reversed s, also < 2^255)
    sc2 = np.ascontiguousarray(sc[::-1]); sd2 = torch.from_numpy(sc2.reshape(-1).copy()).cuda()
    w2 = sc2.view("<u8").astype(object); s2 = w2[:, 0] + (w2[:, 1] << 64) + (w2[:, 2] << 128) + (w2[:, 3] << 192)
    ssum = svals + s2
    sums = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in ssum), dtype=np.uint8) if lg <= 16 else None
    r2 = eng.multiexp_affine(cv.cid, d, sd2, 32, n)
    r1 = eng.multiexp_affine(cv.cid, d, sd, 32, n)
    lhs = eng.normalize(cv.cid, eng.sum_points(cv.cid, r1 + r2, 2))
    total2 = int((ssum * k.astype(object)).sum() % cv.r)
    assert lhs == coracle.normalize(cv.cid, coracle.times_scalar_affine(cv.cid, gen_bytes(cv), total2.to_bytes(32, "little")))
    if sums is not None:
        r3 = eng.multiexp_affine(cv.cid, d, torch.from_numpy(sums.copy()).cuda(), 32, n)
        assert eng.normalize(cv.cid, r3) == lhs


# ---------------------------------------------------------------- "next" row 1: point codecs / batch conversions
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_codecs_match_reference_wasm(eng, cname):
    """g1m_batchLEMtoU / UtoLEM / LEMtoC / CtoLEM / batchToAffine / batchToJacobian against the reference's own exports
    (src/build_curve_jacobian_a0.js:1040-1328,1413-1418), byte for byte."""
    if not refwasm.available(cname): pytest.skip("oracle/_ref not built")
    cv = curve(cname); n8 = cv.n8; n = 257
    pb = refwasm.RefModule(cname)
    bases = bytearray(make_bases(cv, n, 61))
    inf_at = (5, 100, n - 1)
    for i in inf_at: bases[i * 2 * n8:(i + 1) * 2 * n8] = bytes(2 * n8)       # affine infinity
    bases = bytes(bases)

    def ref_batch(fn, data, in_sz, out_sz, cnt):
        mark = pb.heap_mark(); pi = pb.alloc(len(data) + 4 * n8); po = pb.alloc(out_sz * cnt + 64)
        pb.write(pi, data); getattr(pb, fn)(pi, cnt, po); o = pb.read(po, out_sz * cnt); pb.heap_release(mark); return o

    # LEM -> U and back
    u_ref = ref_batch("g1m_batchLEMtoU", bases, 2 * n8, 2 * n8, n)
    u = eng.batch_convert(cv.cid, "LEMtoU", bases, n)
    assert u == u_ref
    assert eng.batch_convert(cv.cid, "UtoLEM", u, n) == ref_batch("g1m_batchUtoLEM", u_ref, 2 * n8, 2 * n8, n) == bases
    # LEM -> C: the reference tests infinity with the Jacobian predicate on an affine input (defect, see codecs.cuh): compare the
    # finite points byte for byte and require the documented 0x40 flag for infinity
    c_ref = ref_batch("g1m_batchLEMtoC", bases, 2 * n8, n8, n)
    c = eng.batch_convert(cv.cid, "LEMtoC", bases, n)
    for i in range(n):
        if i in inf_at: assert c[i * n8] == 0x40 and c[i * n8 + 1:(i + 1) * n8] == bytes(n8 - 1)
        elif (i + 1) in inf_at: assert c_ref[i * n8] == 0x40        # the defect: the reference flags the point BEFORE an all-zero x as infinity
        else: assert c[i * n8:(i + 1) * n8] == c_ref[i * n8:(i + 1) * n8], i
    # C -> LEM (square root + sign selection): ours and the reference's on OUR compressed bytes (the infinity flags are right there)
    back = eng.batch_convert(cv.cid, "CtoLEM", c, n)
    assert back == bases
    assert ref_batch("g1m_batchCtoLEM", c, n8, 2 * n8, n) == bases
    # affine -> Jacobian -> affine; and Jacobian points with z != 1 (MSM partial results)
    j_ref = ref_batch("g1m_batchToJacobian", bases, 2 * n8, 3 * n8, n)
    j = eng.batch_convert(cv.cid, "toJacobian", bases, n)
    assert j == j_ref
    assert eng.batch_convert(cv.cid, "toAffine", j, n) == bases
    sc = make_scalars(64, 9, "u256")
    parts = b"".join(eng.multiexp_affine(cv.cid, bases[k * 8 * 2 * n8:(k + 1) * 8 * 2 * n8], sc[k * 8 * 32:(k + 1) * 8 * 32], 32, 8) for k in range(8))
    parts += bytes(n8) + pyref.fe_bytes(cv, cv.R % cv.q) + bytes(n8)          # g1m_zero
    assert eng.batch_convert(cv.cid, "toAffine", parts, 9) == ref_batch("g1m_batchToAffine", parts, 3 * n8, 2 * n8, 9)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_ffjavascript_style_surface(eng, cname):
    """b200msm.G1: multiExpAffine / multiExp over byte buffers with ffjavascript's argument meaning and error behaviour."""
    import b200msm
    cv = curve(cname); n8 = cv.n8; n = 777
    G = b200msm.G1(eng, cname)
    bases = make_bases(cv, n, 71); sc = make_scalars(n, 72, "u256")
    exp = oracle_msm(cv, bases, sc, 32, n)
    r = G.multiExpAffine(bases, sc)
    assert len(r) == 3 * n8 and eng.normalize(cv.cid, r) == exp
    assert eng.normalize(cv.cid, G.multiExp(G.batchToJacobian(bases), sc)) == exp          # Jacobian bases
    sc16 = b"".join(sc[i * 32:i * 32 + 16] for i in range(n))                               # scalar size is inferred: 16 bytes
    assert eng.normalize(cv.cid, G.multiExpAffine(bases, sc16)) == oracle_msm(cv, bases, sc16, 16, n)
    with pytest.raises(ValueError, match="Scalar size does not match"): G.multiExpAffine(bases, sc[:-1])
    assert G.isZero(G.multiExpAffine(b"", b"")) and G.eq(G.add(r, G.zero()), r)
    assert G.batchUtoLEM(G.batchLEMtoU(bases)) == bases and G.batchCtoLEM(G.batchLEMtoC(bases)) == bases


# ---------------------------------------------------------------- "next" row 2: GLV pre-pass (src/build_glv.js)
GLV_KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "glv.json")))["tests"]


def test_glv_decompose_kat_and_random(eng):
    """g1m_glv_decomposeScalar: the reference's KAT (test/glv.js:50-65), then edge and random 256-bit scalars against
    the restated algorithm (oracle/pyref.py::glv_decompose, pinned on the same KAT) -- exact halves and sign bits."""
    cv = curve("bls12381")
    v = GLV_KAT["decomposeScalar is correct."]["values"]
    k = int(v["scalar"], 16)
    out, signs = eng.glv_decompose_scalars(cv.cid, k.to_bytes(32, "little"), 1)
    assert [int.from_bytes(out[0:32], "little"), int.from_bytes(out[32:64], "little")] == [int(x, 16) for x in v["expectedOutput"]]
    assert signs == [1]
    rnd = random.Random(5)
    ks = [0, 1, 2, cv.r - 1, cv.r, cv.r + 1, 2 * cv.r - 1, 2 * cv.r, 2 * cv.r + 1, (1 << 256) - 1, 1 << 128, (1 << 128) - 1,
          pyref.GLV_LAMBDA, pyref.GLV_U0, pyref.GLV_NEG_V1] + [rnd.getrandbits(256) for _ in range(4000)] + [rnd.getrandbits(rnd.randrange(1, 257)) for _ in range(1000)]
    out, signs = eng.glv_decompose_scalars(cv.cid, b"".join(x.to_bytes(32, "little") for x in ks), len(ks))
    for i, x in enumerate(ks):
        k1, k2, sg = pyref.glv_decompose(x)
        assert (int.from_bytes(out[64 * i:64 * i + 32], "little"), int.from_bytes(out[64 * i + 32:64 * i + 64], "little"), signs[i]) == (k1, k2, sg), hex(x)


def test_glv_preprocess_kat(eng):
    """g1m_glv_preprocessEndomorphism on the reference's own inputs (test/glv.js:103-192): scalar halves match its expected
    output, the points match the restated algorithm, and the MSM over the 2N outputs equals the MSM over the N inputs."""
    cv = curve("bls12381")
    v = GLV_KAT["preprocessEndomorphism is correct."]["values"]
    n = int(v["numPoints"], 16)
    P = [(int(v["inputPoints"][2 * i], 16), int(v["inputPoints"][2 * i + 1], 16)) for i in range(n)]
    S = [int(x, 16) for x in v["inputScalars"][:n]]
    bases = b"".join(pyref.affine_to_bytes(cv, p) for p in P); sc = b"".join(s.to_bytes(32, "little") for s in S)
    p2, s2 = eng.glv_preprocess(cv.cid, bases, sc, n)
    assert [int.from_bytes(s2[32 * i:32 * i + 32], "little") for i in range(2 * n)] == [int(x, 16) for x in v["expectedScalarOutput"][:2 * n]]
    P2, S2 = pyref.glv_preprocess(P, S)
    assert p2 == b"".join(pyref.affine_to_bytes(cv, p) for p in P2)
    assert msm(eng, cv, p2, s2, 32, 2 * n) == oracle_msm(cv, bases, sc, 32, n)


@pytest.mark.parametrize("n", [1, 1000, 1 << 16])
def test_glv_msm_equals_plain_msm(eng, n):
    """test/glv.js:194-252 (the reference's commented-out benchmark check): MSM after the GLV pre-pass == plain MSM, incl. scalars >= r,
    an infinity input and a point with y-negation on both halves.  Also: BN254 is refused like the reference (BLS12-381 only)."""
    cv = curve("bls12381")
    bases = bytearray(make_bases(cv, n, 91)); sc = bytearray(make_scalars(n, 92, "u256"))
    if n >= 1000:
        bases[96 * 7:96 * 8] = bytes(96)                                  # affine infinity
        sc[32 * 9:32 * 10] = (cv.r + 5).to_bytes(32, "little"); sc[32 * 10:32 * 11] = bytes(32)
    bases = bytes(bases); sc = bytes(sc)
    p2, s2 = eng.glv_preprocess(cv.cid, bases, sc, n)
    assert all(s2[32 * i + 16:32 * i + 32] == bytes(16) for i in range(0, 2 * n, max(1, n // 50)))
    plain = msm(eng, cv, bases, sc, 32, n)
    assert msm(eng, cv, p2, s2, 32, 2 * n) == plain
    assert msm(eng, cv, p2, b"".join(s2[32 * i:32 * i + 16] for i in range(2 * n)), 16, 2 * n) == plain   # halves as 16-byte scalars
    if n <= 1000: assert plain == oracle_msm(cv, bases, sc, 32, n)
    import b200msm
    with pytest.raises(b200msm.B200MsmError): eng.glv_preprocess(curve("bn128").cid, bytes(64), bytes(32), 1)


# ---------------------------------------------------------------- resident bases with a precomputed window table
def _resident(eng, cv, h, sc, ssz, n):
    return eng.normalize(cv.cid, eng.multiexp_resident(h, sc, ssz, n, cv.cid))


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("n,wb", [(1, 0), (37, 3), (1000, 0), (1000, 8), (5000, 13), (5000, 16), (3000, 11)])
def test_window_table_matches_oracle(eng, cname, n, wb):
    """b200msm_upload_bases_windowed: all windows share one bucket array fed from the table rows 2^(off_w) * P_i.
    wb = 8 and 16 divide 256 (the unsigned last digit needs the second slot); 3 / 11 / 13 give mixed window widths."""
    cv = curve(cname)
    bases = make_bases(cv, n, 71)
    h = eng.upload_bases_windowed(cv.cid, bases, n, 32, wb)
    try:
        for kind in ("u256", "equal", "small"):
            sc = make_scalars(n, 72, kind)
            assert _resident(eng, cv, h, sc, 32, n) == oracle_msm(cv, bases, sc, 32, n), kind
        assert _resident(eng, cv, h, bytes(32 * n), 32, n) == bytes(2 * cv.n8)               # all-zero scalars
        sc = b"".join(v.to_bytes(32, "little") for v in ([cv.r - 1, cv.r, (1 << 256) - 1, 1 << 255, 1, 0, (1 << 128) - 1] * n)[:n])
        assert _resident(eng, cv, h, sc, 32, n) == oracle_msm(cv, bases, sc, 32, n)
        if n > 100:
            m = n - 33                                                                        # fewer points than uploaded
            sc = make_scalars(m, 73, "u256")
            assert _resident(eng, cv, h, sc, 32, m) == oracle_msm(cv, bases[:m * 2 * cv.n8], sc, 32, m)
            sc16 = bytes(random.Random(74).getrandbits(8) for _ in range(16 * n))             # other scalar size: ordinary pipeline, same handle
            assert _resident(eng, cv, h, sc16, 16, n) == oracle_msm(cv, bases, sc16, 16, n)
            eng.set_option("combine", 1)
            try:
                sc = make_scalars(n, 75, "u256")
                assert _resident(eng, cv, h, sc, 32, n) == oracle_msm(cv, bases, sc, 32, n)
            finally:
                eng.set_option("combine", 0)
    finally:
        eng.free_bases(h)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_window_table_special_points_and_lanes(eng, cname):
    """duplicate points, P and -P, infinity inputs (their table rows stay infinity), any lane / group split, 16-byte scalars"""
    cv = curve(cname); n8 = cv.n8
    g = gen_bytes(cv); ng = pyref.affine_to_bytes(cv, pyref.neg(cv, cv.G)); zero = bytes(2 * n8)
    pts = (g + g + ng + zero + g + ng + ng + zero) * 40 + make_bases(cv, 3000, 81)
    m = len(pts) // (2 * n8)
    h = eng.upload_bases_windowed(cv.cid, pts, m, 32, 10)
    h16 = eng.upload_bases_windowed(cv.cid, pts, m, 16, 0)
    try:
        for lanes, gp in ((1, 0), (4, 0), (3, 7000), (2, 1500)):
            eng.set_option("lanes", lanes); eng.set_option("group_pairs", gp)
            for seed, kind in ((1, "equal"), (2, "small"), (3, "u256")):
                sc = make_scalars(m, seed, kind)
                assert _resident(eng, cv, h, sc, 32, m) == oracle_msm(cv, pts, sc, 32, m), (lanes, gp, kind)
        sc16 = bytes(random.Random(82).getrandbits(8) for _ in range(16 * m))
        assert _resident(eng, cv, h16, sc16, 16, m) == oracle_msm(cv, pts, sc16, 16, m)
    finally:
        eng.set_option("lanes", 4); eng.set_option("group_pairs", 0)
        eng.free_bases(h); eng.free_bases(h16)


@pytest.mark.parametrize("cname,lg", [("bls12381", 18), ("bn128", 18)])
def test_window_table_full_size_known_answer(eng, cname, lg):
    """2^18 points through the window table against the exact known answer (sum_i s_i k_i mod r) * G and against the ordinary path"""
    import numpy as np, torch
    cv = curve(cname); n = 1 << lg; n8 = cv.n8
    seed = 0xB2000000 + lg
    d = torch.empty(n * 2 * n8, dtype=torch.uint8, device="cuda")
    eng.generate_bases(cv.cid, seed, 0, n, d)
    with np.errstate(over="ignore"):
        k = _splitmix64_np(np.uint64(seed) + np.arange(n, dtype=np.uint64))
    k[k == 0] = 1
    sc = np.random.default_rng(lg + 100).integers(0, 256, size=(n, 32), dtype=np.uint8)
    sd = torch.from_numpy(sc.reshape(-1).copy()).cuda()
    words = sc.view("<u8").astype(object)
    svals = words[:, 0] + (words[:, 1] << 64) + (words[:, 2] << 128) + (words[:, 3] << 192)
    total = int((svals * k.astype(object)).sum() % cv.r)
    exp = coracle.normalize(cv.cid, coracle.times_scalar_affine(cv.cid, gen_bytes(cv), total.to_bytes(32, "little")))
    h = eng.upload_bases_windowed(cv.cid, d, n, 32, 0)
    try:
        assert _resident(eng, cv, h, sd, 32, n) == exp
        assert msm(eng, cv, d, sd, 32, n) == exp
    finally:
        eng.free_bases(h)


# ---------------------------------------------------------------- batched MSMs over one resident base set (BASELINE config 5)
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("windowed", [False, True])
def test_batch_matches_individual_msms(eng, cname, windowed):
    """b200msm_g1_multiexp_batch: every MSM of the batch equals the oracle's result for its scalar block, for any worker count;
    device and host buffers."""
    import torch
    cv = curve(cname); n = 700; count = 11
    bases = make_bases(cv, n, 95)
    sc = b"".join(make_scalars(n, 200 + j, "u256") for j in range(count))
    exp = [oracle_msm(cv, bases, sc[j * n * 32:(j + 1) * n * 32], 32, n) for j in range(count)]
    h = eng.upload_bases_windowed(cv.cid, bases, n, 32, 0) if windowed else eng.upload_bases(cv.cid, bases, n)
    try:
        for workers in (1, 3, 4):
            eng.set_option("batch_workers", workers)
            out = eng.multiexp_batch(h, sc, 32, n, count, cv.cid)
            sz = 3 * cv.n8
            assert [eng.normalize(cv.cid, out[j * sz:(j + 1) * sz]) for j in range(count)] == exp, workers
        import numpy as np
        sd = torch.from_numpy(np.frombuffer(sc, dtype=np.uint8).copy()).cuda(); od = torch.zeros(count * 3 * cv.n8, dtype=torch.uint8, device="cuda")
        eng.multiexp_batch(h, sd, 32, n, count, cv.cid, out=od)
        torch.cuda.synchronize()
        assert eng.normalize(cv.cid, od, count) == b"".join(exp)
    finally:
        eng.set_option("batch_workers", 4); eng.free_bases(h)


# ---------------------------------------------------------------- "next" row 3: G2 (g2m_* over Fq2) through the same entry points
G2ID = {"bls12381": 2, "bn128": 3}
_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


@pytest.fixture(scope="module", params=["bls12381", "bn128"])
def g2ref(request):
    if not refwasm.available(request.param): pytest.skip("oracle/_ref not built")
    return request.param, refwasm.RefG2(refwasm.RefModule(request.param))


def _g2_bases(ref, n, seed):
    """P_i = k_i * G2 from the reference's own g2m_timesScalarAffine (k_i = splitmix64(seed + i), the engine's stream)"""
    G = ref.generator_affine()
    return b"".join(ref.times_scalar_affine(G, (_splitmix64(seed + i) or 1).to_bytes(8, "little")) for i in range(n))


def _g2_msm(eng, cid, bases, sc, ssz, n):
    return eng.normalize(cid, eng.multiexp_affine(cid, bases, sc, ssz, n))


def test_g2_fq2_ops_match_reference(eng, g2ref):
    """f2m_mul / add / sub / square / inverse / neg / toMontgomery / fromMontgomery (build_f2m.js) through b200msm_fq_op"""
    cname, ref = g2ref; cv = curve(cname); cid = G2ID[cname]; e8 = 2 * cv.n8
    rnd = random.Random(41)
    def el(c0, c1): return c0.to_bytes(cv.n8, "little") + c1.to_bytes(cv.n8, "little")
    vals = [el(0, 0), el(1, 0), el(0, 1), el(cv.q - 1, cv.q - 1), el(cv.R % cv.q, 0), el(5, 0), el(0, 7)] + [el(rnd.randrange(cv.q), rnd.randrange(cv.q)) for _ in range(60)]
    a = b"".join(vals); b = b"".join(reversed(vals)); n = len(vals)
    for op, fn, two in ((0, "mul", True), (1, "add", True), (2, "sub", True), (3, "square", False), (4, "inverse", False), (5, "toMontgomery", False),
                        (6, "fromMontgomery", False), (7, "neg", False), (8, "inverse", False)):
        got = eng.fq_op(cid, op, a, b if two else None)
        for i in range(n):
            x = a[i * e8:(i + 1) * e8]; y = b[i * e8:(i + 1) * e8]
            if fn == "inverse" and x == bytes(e8): continue
            assert got[i * e8:(i + 1) * e8] == ref.f2m(fn, x, y if two else None), (fn, i)


def test_g2_generate_bases_matches_reference(eng, g2ref):
    import torch
    cname, ref = g2ref; cv = curve(cname); cid = G2ID[cname]; n = 40
    d = torch.empty(n * 4 * cv.n8, dtype=torch.uint8, device="cuda")
    eng.generate_bases(cid, 777, 5, n, d)
    assert bytes(d.cpu().numpy()) == _g2_bases(ref, n, 777 + 5)


@pytest.mark.parametrize("n", [1, 2, 3, 10, 100, 1000])
def test_g2_msm_matches_reference_wasm(eng, g2ref, n):
    """g2m_multiexpAffine of the reference module vs the engine, canonical affine output, uniform 256-bit scalars"""
    cname, ref = g2ref; cv = curve(cname); cid = G2ID[cname]
    bases = _g2_bases(ref, min(n, 120), 900 + n) * (n // 120 + 1)
    bases = bases[: n * 4 * cv.n8]
    sc = make_scalars(n, 901 + n, "u256")
    assert _g2_msm(eng, cid, bases, sc, 32, n) == ref.msm_affine(bases, sc, 32, n)


def test_g2_edge_cases_chunks_and_variants(eng, g2ref):
    cname, ref = g2ref; cv = curve(cname); cid = G2ID[cname]; pt = 4 * cv.n8
    zero = bytes(pt)
    assert _g2_msm(eng, cid, b"", b"", 32, 0) == zero
    G = ref.generator_affine()
    Gp = pyref.g2_from_bytes(cv, G); R = cv.R
    negG = b"".join(((c * R) % cv.q).to_bytes(cv.n8, "little") for c in (Gp[0][0], Gp[0][1], (-Gp[1][0]) % cv.q, (-Gp[1][1]) % cv.q))
    five = (5).to_bytes(32, "little")
    assert _g2_msm(eng, cid, G + negG, five * 2, 32, 2) == zero
    base = _g2_bases(ref, 60, 33)
    pts = (G + G + negG + zero + G + negG + negG + zero) * 10 + base
    m = len(pts) // pt
    for seed, kind in ((1, "equal"), (2, "small"), (3, "u256")):
        sc = make_scalars(m, seed, kind)
        assert _g2_msm(eng, cid, pts, sc, 32, m) == ref.msm_affine(pts, sc, 32, m), kind
    assert _g2_msm(eng, cid, pts, bytes(32 * m), 32, m) == zero
    special = [cv.r - 1, cv.r, cv.r + 1, (1 << 256) - 1, 1 << 255, 1, 0, (1 << 128) - 1]
    sc = b"".join(v.to_bytes(32, "little") for v in special)
    assert _g2_msm(eng, cid, base[: 8 * pt], sc, 32, 8) == ref.msm_affine(base[: 8 * pt], sc, 32, 8)
    # per-window export and other scalar sizes
    sc = make_scalars(60, 7, "u256")
    for start, bits in ((0, 5), (13, 7), (250, 11)):
        got = eng.normalize(cid, eng.multiexp_affine_chunk(cid, base, sc, 32, 60, start, bits))
        assert got == ref.msm_chunk(base, sc, 32, 60, start, bits), (start, bits)
    sc5 = bytes(random.Random(8).getrandbits(8) for _ in range(5 * 60))
    assert _g2_msm(eng, cid, base, sc5, 5, 60) == ref.msm_affine(base, sc5, 5, 60)
    # window widths, device window chain, window table, batch, sum of partials
    sc = make_scalars(60, 9, "u256"); exp = ref.msm_affine(base, sc, 32, 60)
    try:
        for wb in (1, 4, 8, 13):
            eng.set_option("window_bits", wb)
            assert _g2_msm(eng, cid, base, sc, 32, 60) == exp, wb
        eng.set_option("window_bits", 0); eng.set_option("combine", 1)
        assert _g2_msm(eng, cid, base, sc, 32, 60) == exp
    finally:
        eng.set_option("window_bits", 0); eng.set_option("combine", 0)
    for wb in (0, 8, 11):
        h = eng.upload_bases_windowed(cid, base, 60, 32, wb)
        try:
            assert eng.normalize(cid, eng.multiexp_resident(h, sc, 32, 60, cid)) == exp, wb
            out = eng.multiexp_batch(h, sc + make_scalars(60, 10, "u256"), 32, 60, 2, cid)
            assert eng.normalize(cid, out[: 3 * 2 * cv.n8]) == exp
        finally:
            eng.free_bases(h)
    a = eng.multiexp_affine(cid, base[: 30 * pt], sc[: 30 * 32], 32, 30); b = eng.multiexp_affine(cid, base[30 * pt:], sc[30 * 32:], 32, 30)
    assert eng.normalize(cid, eng.sum_points(cid, a + b, 2)) == exp


@pytest.mark.parametrize("lg", [14, 16])
def test_g2_full_size_known_answer(eng, g2ref, lg):
    """P_i = k_i * G2 generated on the device (stream checked above): sum_i s_i P_i = (sum_i s_i k_i mod r) * G2, the right-hand
    side from ONE g2m_timesScalarAffine of the reference; ordinary path and window table."""
    import numpy as np, torch
    cname, ref = g2ref; cv = curve(cname); cid = G2ID[cname]; n = 1 << lg
    seed = 0xB2000000 + lg
    d = torch.empty(n * 4 * cv.n8, dtype=torch.uint8, device="cuda")
    eng.generate_bases(cid, seed, 0, n, d)
    with np.errstate(over="ignore"):
        k = _splitmix64_np(np.uint64(seed) + np.arange(n, dtype=np.uint64))
    k[k == 0] = 1
    sc = np.random.default_rng(lg).integers(0, 256, size=(n, 32), dtype=np.uint8)
    sd = torch.from_numpy(sc.reshape(-1).copy()).cuda()
    words = sc.view("<u8").astype(object)
    svals = words[:, 0] + (words[:, 1] << 64) + (words[:, 2] << 128) + (words[:, 3] << 192)
    total = int((svals * k.astype(object)).sum() % cv.r)
    m = ref.m; mark = m.heap_mark(); pB = m.alloc(4 * cv.n8); pS = m.alloc(40); pR = m.alloc(6 * cv.n8)
    m.write(pB, ref.generator_affine()); m.write(pS, total.to_bytes(32, "little")); m.g2m_timesScalarAffine(pB, pS, 32, pR)
    exp = ref.canonical(pR); m.heap_release(mark)
    assert _g2_msm(eng, cid, d, sd, 32, n) == exp
    h = eng.upload_bases_windowed(cid, d, n, 32, 0)
    try:
        assert eng.normalize(cid, eng.multiexp_resident(h, sd, 32, n, cid)) == exp
    finally:
        eng.free_bases(h)


# ---------------------------------------------------------------- "next" row 4: Fr NTT (frm_fft / frm_ifft, src/build_fft.js)
def _fr_elems(cv, n, seed):
    rnd = random.Random(seed); R = 1 << 256
    return b"".join((rnd.randrange(cv.r) * R % cv.r).to_bytes(32, "little") for _ in range(n))


def _ref_fft(pb, data, n, inverse):
    mark = pb.heap_mark(); p = pb.alloc(len(data) + 64); pb.write(p, data)
    (pb.frm_ifft if inverse else pb.frm_fft)(p, n)
    out = pb.read(p, len(data)); pb.heap_release(mark); return out


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_fr_fft_matches_reference_wasm(eng, cname):
    """frm_fft / frm_ifft of the reference module vs b200msm_fr_fft for every size 2^0 .. 2^13 (tile-only, radix-4 and radix-2 tails),
    byte for byte on Montgomery Fr elements; plus 2^16."""
    if not refwasm.available(cname): pytest.skip("oracle/_ref not built")
    cv = curve(cname); pb = refwasm.RefModule(cname)
    for lg in list(range(0, 14)) + [16]:
        n = 1 << lg
        data = _fr_elems(cv, n, 500 + lg)
        if lg == 3: data = bytes(32) * 3 + data[96:]                       # zeros among the inputs
        for inv in (False, True):
            assert eng.fr_fft(cv.cid, data, lg, inverse=inv) == _ref_fft(pb, data, n, inv), (lg, inv)


@pytest.mark.parametrize("cname,lg", [("bls12381", 20), ("bn128", 22)])
def test_fr_fft_large_properties(eng, cname, lg):
    """size-independent checks at sizes the CPU reference does not finish quickly: ifft(fft(x)) == x in place on the device,
    linearity, out[0] = sum of the inputs, and a delta at position 1 transforms to the powers of the root."""
    import numpy as np, torch
    cv = curve(cname); n = 1 << lg; R = 1 << 256
    rng = np.random.default_rng(lg)
    x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); x[:, 31] &= 0x0F          # < 2^252 < r: valid reduced elements
    y = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); y[:, 31] &= 0x0F
    xd = torch.from_numpy(x.reshape(-1).copy()).cuda(); yd = torch.from_numpy(y.reshape(-1).copy()).cuda()
    fx = torch.empty_like(xd); fy = torch.empty_like(yd)
    torch.cuda.synchronize()
    eng.fr_fft(cv.cid, xd, lg, out=fx); eng.fr_fft(cv.cid, yd, lg, out=fy)
    eng.synchronize()                                                                # the engine runs on its own stream
    back = fx.clone(); torch.cuda.synchronize()
    eng.fr_fft(cv.cid, back, lg, inverse=True, out=back)                             # in place
    eng.synchronize()
    assert torch.equal(back, xd)
    # linearity on a sample of positions: F(x + y)[k] == F(x)[k] + F(y)[k]  (field addition via the engine's Fr... done on the host)
    def ints(t, idx): b = bytes(t[idx * 32:(idx + 1) * 32].cpu().numpy()); return int.from_bytes(b, "little")
    xs = [int.from_bytes(x[i].tobytes(), "little") for i in range(n)] if lg <= 20 else None
    z = (x.astype(np.uint16) + y.astype(np.uint16))                                   # byte-wise sums with carries resolved below
    carry = np.zeros(n, dtype=np.uint16); zb = np.zeros((n, 32), dtype=np.uint8)
    for k in range(32):
        t = z[:, k] + carry; zb[:, k] = (t & 0xFF).astype(np.uint8); carry = t >> 8
    assert int(carry.max()) == 0                                                      # x + y < 2^253 < r: still reduced
    zd = torch.from_numpy(zb.reshape(-1).copy()).cuda(); fz = torch.empty_like(zd)
    torch.cuda.synchronize()
    eng.fr_fft(cv.cid, zd, lg, out=fz); eng.synchronize()
    for k in (0, 1, 2, n // 2, n - 1, 12345 % n):
        assert ints(fz, k) == (ints(fx, k) + ints(fy, k)) % cv.r, k
    if xs is not None: assert ints(fx, 0) == sum(xs) % cv.r
    # delta at position 1 (Montgomery one) -> out[k] = w^k: out[1]^(n) == 1 and out[2] == out[1]^2 (as field elements)
    d = np.zeros((n, 32), dtype=np.uint8); d[1] = np.frombuffer((R % cv.r).to_bytes(32, "little"), dtype=np.uint8)
    fd = bytes(eng.fr_fft(cv.cid, d.tobytes(), lg)[: 4 * 32])
    Ri = pow(R, -1, cv.r)
    w0, w1, w2, w3 = [int.from_bytes(fd[i * 32:(i + 1) * 32], "little") * Ri % cv.r for i in range(4)]
    assert w0 == 1 and w2 == w1 * w1 % cv.r and w3 == w2 * w1 % cv.r and pow(w1, n, cv.r) == 1 and pow(w1, n // 2, cv.r) == cv.r - 1


def test_ffjavascript_style_g2_surface(eng, g2ref):
    """b200msm.G2: curve.G2.multiExpAffine / multiExp / add / eq / zero over byte buffers"""
    import b200msm
    cname, ref = g2ref; cv = curve(cname); e8 = 2 * cv.n8; n = 50
    G2 = b200msm.G2(eng, cname)
    bases = _g2_bases(ref, n, 301); sc = make_scalars(n, 302, "u256")
    exp = ref.msm_affine(bases, sc, 32, n)
    r = G2.multiExpAffine(bases, sc)
    assert len(r) == 3 * e8 and eng.normalize(G2.curve, r) == exp
    one2 = (cv.R % cv.q).to_bytes(cv.n8, "little") + bytes(cv.n8)
    jac = b"".join(bases[i * 2 * e8:(i + 1) * 2 * e8] + one2 for i in range(n))
    assert eng.normalize(G2.curve, G2.multiExp(jac, sc)) == exp
    assert G2.isZero(G2.multiExpAffine(b"", b"")) and G2.eq(G2.add(r, G2.zero()), r)
    with pytest.raises(ValueError, match="Base size does not match"): G2.multiExpAffine(bases[:-1], sc)


# ---------------------------------------------------------------- Jacobian bases: g1m_multiexp / g2m_multiexp (n8b = 3*n8)
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_jacobian_bases_match_reference_wasm(eng, cname):
    """b200msm_g1_multiexp[_chunk] vs the reference module's g1m_multiexp / g2m_multiexp on bases (X, Y, Z) with random Z, Z = 1 and Z = 0"""
    if not refwasm.available(cname): pytest.skip("oracle/_ref not built")
    cv = curve(cname); n8 = cv.n8; q = cv.q; R = cv.R; rnd = random.Random(77)
    pb = refwasm.RefModule(cname)
    n = 300
    aff = make_bases(cv, n, 123)
    jac = bytearray()
    for i in range(n):
        x = int.from_bytes(aff[i * 2 * n8: i * 2 * n8 + n8], "little"); y = int.from_bytes(aff[i * 2 * n8 + n8: (i + 1) * 2 * n8], "little")   # Montgomery
        if i % 7 == 3: jac += bytes(3 * n8); continue                                             # infinity (z = 0)
        z = 1 if i % 5 == 0 else rnd.randrange(1, q)
        Ri = pow(R, -1, q); xs = x * Ri % q; ys = y * Ri % q
        for c in (xs * z * z % q, ys * z * z * z % q, z): jac += (c * R % q).to_bytes(n8, "little")
    jac = bytes(jac); sc = make_scalars(n, 124, "u256")
    want = pb.normalize_bytes(refwasm.msm_jacobian(pb, "g1m", jac, sc, 32, n))
    assert eng.normalize(cv.cid, eng.multiexp_jacobian(cv.cid, jac, sc, 32, n)) == pyref.canonical_bytes(cv, want)
    for ch in ((0, 7), (100, 13), (250, 11)):
        want = pb.normalize_bytes(refwasm.msm_jacobian(pb, "g1m", jac, sc, 32, n, ch))
        assert eng.normalize(cv.cid, eng.multiexp_jacobian(cv.cid, jac, sc, 32, n, ch)) == pyref.canonical_bytes(cv, want), ch
    assert eng.normalize(cv.cid, eng.multiexp_jacobian(cv.cid, b"", b"", 32, 0)) == bytes(2 * n8)
    # G2: bases k_i * G2 with Z = 1 (and two infinities)
    g2 = refwasm.RefG2(pb); cid2 = G2ID[cname]; e8 = 2 * n8; m = 40
    base = _g2_bases(g2, m, 55)
    one2 = (R % q).to_bytes(n8, "little") + bytes(n8)
    jac2 = b"".join((bytes(3 * e8) if i in (4, 17) else base[i * 2 * e8:(i + 1) * 2 * e8] + one2) for i in range(m))
    sc2 = make_scalars(m, 56, "u256")
    want2 = g2.canonical_of(refwasm.msm_jacobian(pb, "g2m", jac2, sc2, 32, m))
    assert eng.normalize(cid2, eng.multiexp_jacobian(cid2, jac2, sc2, 32, m)) == want2


# ---------------------------------------------------------------- the reference's own test bodies through the Protoboard mirror
def test_reference_fft_and_glv_tests_through_protoboard(eng):
    """test/fft.js:14-33 ("fft / ifft of N values gives the values back") and test/glv.js:50-65, 103-192 written against
    b200msm.Protoboard exactly as the reference writes them against its WASM protoboard."""
    import b200msm
    pb = b200msm.Protoboard("bls12381", engine=eng)
    n8r = 32; n8q = 48
    # --- test/fft.js: N = 1024 values i+1 in Montgomery form, fft then ifft, compare
    N = 1024
    p = pb.alloc(n8r * N)
    r = pyref.BLS12_381.r; R = 1 << 256
    for i in range(N): pb.set(p + i * n8r, (i + 1) * R % r, n8r)           # frm_toMontgomery(i + 1)
    before = pb.read(p, n8r * N)
    pb.frm_fft(p, N)
    assert pb.read(p, n8r * N) != before
    pb.frm_ifft(p, N)
    assert pb.read(p, n8r * N) == before
    # --- test/glv.js:50-65
    v = GLV_KAT["decomposeScalar is correct."]["values"]
    pScalar = pb.alloc(32); pb.set(pScalar, int(v["scalar"], 16), 32)
    pScalarRes = pb.alloc(64)
    sign = pb.g1m_glv_decomposeScalar(pScalar, pScalarRes)
    assert pb.get(pScalarRes, 2, 32) == [int(x, 16) for x in v["expectedOutput"]] and sign == 1
    # --- test/glv.js:103-192
    v = GLV_KAT["preprocessEndomorphism is correct."]["values"]
    numPoints = int(v["numPoints"], 16)
    pPoints = pb.alloc(numPoints * n8q * 2); pScalars = pb.alloc(numPoints * n8r)
    pPre = pb.alloc(numPoints * n8q * 4); pPreS = pb.alloc(numPoints * n8r * 2); pRes = pb.alloc(n8q * 3); pExp = pb.alloc(n8q * 3)
    for i in range(numPoints):
        pb.set(pPoints + 96 * i, int(v["inputPoints"][2 * i], 16), 48); pb.set(pPoints + 96 * i + 48, int(v["inputPoints"][2 * i + 1], 16), 48)
        pb.f1m_toMontgomery(pPoints + 96 * i, pPoints + 96 * i); pb.f1m_toMontgomery(pPoints + 96 * i + 48, pPoints + 96 * i + 48)
        pb.set(pScalars + n8r * i, int(v["inputScalars"][i], 16), n8r)
    pb.g1m_glv_preprocessEndomorphism(pPoints, pScalars, numPoints, pPre, pPreS)
    pb.g1m_multiexp_multiExp(pPre, pPreS, numPoints * 2, pRes)
    pb.g1m_normalize(pRes, pRes)
    assert pb.get(pPreS, numPoints * 2, n8r) == [int(x, 16) for x in v["expectedScalarOutput"][: 2 * numPoints]]
    pb.g1m_multiexpAffine(pPoints, pScalars, n8r, numPoints, pExp)
    pb.g1m_normalize(pExp, pExp)
    assert pb.get(pRes, 2, 48) == pb.get(pExp, 2, 48)
