"""pyref.py -- TEST INFRASTRUCTURE (oracle), not product code.

Pure-Python big-integer restatement of the G1 MSM path of the reference
(Manta-Network/zprize-wasm-msm, wasmcurves fork).  Slow, obviously-correct; used
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as a
checker.  All citations are relative to /root/reference/wasmcurves/.

Pinned against: the reference's golden vectors (tests/golden/*.json, extracted
from test/batchAffine.js, test/glv.js, test/bn128.js, test/bls12381.js) and
against the reference's own compiled WASM run natively (oracle/_ref, see
oracle/refwasm.py) -- tests/test_oracle.py.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Curve:
    name: str
    cid: int          # matches include/b200msm.h: B200MSM_BLS12_381_G1 = 0, B200MSM_BN254_G1 = 1
    q: int            # base field modulus   (src/bls12381/build_bls12381.js:22, src/bn128/build_bn128.js:20)
    r: int            # group order          (build_bls12381.js:23, build_bn128.js:21)
    b: int            # y^2 = x^3 + b
    gx: int
    gy: int
    n8: int           # bytes per Fq element (48 / 32)

    @property
    def n32(self): return self.n8 // 4

    @property
    def R(self): return 1 << (8 * self.n8)      # Montgomery radix, src/build_f1m.js:30-32,42-43

    @property
    def G(self): return (self.gx, self.gy)


BLS12_381 = Curve(
    "bls12381", 0,
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    4,
    3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
    1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569,
    48)

BN254 = Curve(
    "bn128", 1,
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    3, 1, 2, 32)

CURVES = {"bls12381": BLS12_381, "bn128": BN254, 0: BLS12_381, 1: BN254}

# ---------------------------------------------------------------- field / encoding


def to_mont(cv, a): return (a * cv.R) % cv.q          # f1m_toMontgomery  build_f1m.js:1089


def from_mont(cv, a): return (a * pow(cv.R, -1, cv.q)) % cv.q   # f1m_fromMontgomery build_f1m.js:1098


def fe_bytes(cv, a): return int(a).to_bytes(cv.n8, "little")


def fe_from(cv, b): return int.from_bytes(b, "little")


# ---------------------------------------------------------------- affine group law (None = infinity)

def is_on_curve(cv, P):
    if P is None: return True
    x, y = P
    return (y * y - x * x * x - cv.b) % cv.q == 0


def neg(cv, P):
    return None if P is None else (P[0], (-P[1]) % cv.q)


def add(cv, P, Q):
    """Complete affine addition.  The reference's batch-affine formulas
    (src/build_multiexp_opt.js:1090-1239) are the generic branch of this; the
    P+(-P) and infinity branches are where the reference is defective."""
    if P is None: return Q
    if Q is None: return P
    q = cv.q
    x1, y1 = P; x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % q == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, q) % q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, q) % q
    x3 = (lam * lam - x1 - x2) % q
    y3 = (lam * (x1 - x3) - y1) % q
    return (x3, y3)


def mul(cv, k, P):
    """k*P, double-and-add (stands in for g1m_timesScalar, src/build_timesscalarnaf.js)."""
    R = None
    for bit in bin(k)[2:] if k else "":
        R = add(cv, R, R)
        if bit == "1": R = add(cv, R, P)
    return R


def msm_naive(cv, points, scalars):
    acc = None
    for P, k in zip(points, scalars):
        acc = add(cv, acc, mul(cv, k, P))
    return acc


# ---------------------------------------------------------------- byte layouts at the boundary

def affine_to_bytes(cv, P):
    """affine Montgomery LE x||y; infinity = all-zero (g1m_isZeroAffine, build_curve_jacobian_a0.js:55-77)"""
    if P is None: return bytes(2 * cv.n8)
    return fe_bytes(cv, to_mont(cv, P[0])) + fe_bytes(cv, to_mont(cv, P[1]))


def affine_from_bytes(cv, b):
    x = fe_from(cv, b[:cv.n8]); y = fe_from(cv, b[cv.n8:2 * cv.n8])
    if x == 0 and y == 0: return None
    return (from_mont(cv, x), from_mont(cv, y))


def jacobian_from_bytes(cv, b):
    """Jacobian Montgomery x||y||z -> canonical affine (g1m_normalize + f1m_fromMontgomery,
    build_curve_jacobian_a0.js:940-973; what test/batchAffine.js:1249-1254 compares)."""
    n8 = cv.n8
    X, Y, Z = (from_mont(cv, fe_from(cv, b[i * n8:(i + 1) * n8])) for i in range(3))
    if Z == 0: return None
    zi = pow(Z, -1, cv.q)
    return (X * zi * zi % cv.q, Y * zi * zi * zi % cv.q)


def canonical_bytes(cv, P):
    """x||y as plain (non-Montgomery) LE integers < q, infinity = zeros: the parity format."""
    if P is None: return bytes(2 * cv.n8)
    return fe_bytes(cv, P[0]) + fe_bytes(cv, P[1])


# ---------------------------------------------------------------- upstream MSM (src/build_multiexp.js)

WASMCURVE_TSIZES = [17, 17, 17, 17, 17, 17, 17, 17, 17, 17, 16, 16, 15, 14, 13, 13,
                    12, 11, 10, 9, 8, 7, 7, 6, 5, 4, 3, 2, 1, 1, 1, 1]      # build_multiexp.js:275-280


def clz32(n): return 32 - n.bit_length()


def get_chunk(scalar_bytes, scalar_size, start_bit, chunk_size):
    """_getChunk, build_multiexp.js:25-94 (== build_multiexp_opt.js:1251-1322).
    Loads an unaligned u32 at byte startBit/8 (over-reading past the scalar end is
    masked away by bitsToEnd; we pad with zeros), shifts, masks."""
    bits_to_end = scalar_size * 8 - start_bit
    mask = (1 << (bits_to_end if chunk_size > bits_to_end else chunk_size)) - 1
    o = start_bit >> 3
    w = int.from_bytes((scalar_bytes[o:o + 4] + b"\0\0\0\0")[:4], "little")
    return (w >> (start_bit & 7)) & mask


def reduce_table(cv, table, p):
    """_reduceTable, build_multiexp.js:373-461: in-place recursive halving so that
    table[0] ends as sum_{k} (k+1)*table[k]."""
    if p == 1: return
    half = 1 << (p - 1)
    acc_i = half - 1
    for i in range(half - 1):
        table[i] = add(cv, table[i], table[half + i])
        table[acc_i] = add(cv, table[acc_i], table[half + i])
    reduce_table(cv, table, p - 1)
    for _ in range(p - 1):
        table[acc_i] = add(cv, table[acc_i], table[acc_i])
    table[0] = add(cv, table[0], table[acc_i])


def multiexp_chunk(cv, points, scalars_bytes, scalar_size, n, start_bit, chunk_size):
    """g1m_multiexpAffine_chunk, build_multiexp.js:96-249: sum_i digit_i * P_i, no 2^startBit factor."""
    if n == 0: return None
    table = [None] * (1 << chunk_size)
    for i in range(n):
        idx = get_chunk(scalars_bytes[i * scalar_size:(i + 1) * scalar_size], scalar_size, start_bit, chunk_size)
        if idx: table[idx - 1] = add(cv, table[idx - 1], points[i])
    reduce_table(cv, table, chunk_size)
    return table[0]


def multiexp_affine(cv, points, scalars_bytes, scalar_size, n):
    """g1m_multiexpAffine, build_multiexp.js:251-371."""
    pr = None
    if n == 0: return pr
    c = WASMCURVE_TSIZES[clz32(n)]
    n_chunks = (scalar_size * 8 - 1) // c + 1
    it_bit = (n_chunks - 1) * c
    while it_bit >= 0:
        if pr is not None:
            for _ in range(c): pr = add(cv, pr, pr)
        pr = add(cv, pr, multiexp_chunk(cv, points, scalars_bytes, scalar_size, n, it_bit, c))
        it_bit -= c
    return pr


# ---------------------------------------------------------------- Manta opt path: schedule pieces (KAT-able)

def opt_bucket_width(n):
    """_getOptimalBucketWidth, build_multiexp_opt.js:33-49, table :39-44 indexed by clz32(n)."""
    t = [17, 17, 17, 17, 17, 17, 17, 17, 17, 17, 16, 16, 14, 13, 12, 12,
         11, 11, 10, 9, 8, 7, 7, 6, 5, 4, 3, 2, 1, 1, 1, 1]
    return t[clz32(n)]


SENTINEL = 0xFFFFFFFFFFFFFFFF


def compute_schedule(scalars_bytes, n, c, scalar_size=32):
    """_computeSchedule / _singlePointComputeSchedule, build_multiexp_opt.js:175-347.
    schedule[k*n+i] = (i<<32)|digit, or SENTINEL when digit==0; round_counts[k] = #non-zero."""
    W = (scalar_size * 8 + c - 1) // c
    sched = [SENTINEL] * (W * n); counts = [0] * W
    for i in range(n):
        s = scalars_bytes[i * scalar_size:(i + 1) * scalar_size]
        for k in range(W):
            d = get_chunk(s, scalar_size, k * c, c)
            if d:
                sched[k * n + i] = (i << 32) | d; counts[k] += 1
    return sched, counts


def organize_buckets_one_round(sched_k, c):
    """_organizeBucketsOneRound, build_multiexp_opt.js:364-561: stable counting sort by digit,
    sentinels moved to the tail; returns (sorted, bucket_counts[2^c])."""
    cnt = [0] * (1 << c)
    for w in sched_k:
        if w != SENTINEL: cnt[w & 0x7FFFFFFF] += 1
    live = sorted((w for w in sched_k if w != SENTINEL), key=lambda w: w & 0x7FFFFFFF)
    return live + [SENTINEL] * (len(sched_k) - len(live)), cnt


# ---------------------------------------------------------------- GLV (src/build_glv.js) -- "next" row, kept for KATs

BLS_Z = 0xd201000000010000
BLS_LAMBDA = BLS12_381.r - BLS_Z * BLS_Z        # lambda with phi(P) = lambda*P, SURVEY section 4


# GLV constants and decomposition, restated from src/build_glv.js:13-22 and :53-146 (BLS12-381 only, like the reference)
GLV_NEG_V1 = 228988810152649578064853576960394133503          # -v1   (build_glv.js:19)
GLV_U0 = 228988810152649578064853576960394133504              # u0    (build_glv.js:17); u1 = v0 = 1
GLV_BETA = 793479390729215512621379701633421447060886740281060493010456487427281649075476305620758731620350   # :21
GLV_DIVISOR = BLS12_381.r                                      # v0*u1 - v1*u0 = r (build_glv.js:22)


def glv_decompose(k):
    """g1m_glv_decomposeScalar (build_glv.js:53-146): returns (|k1| mod 2^128, |k2| mod 2^128, sign) with
    sign bit 0 = (k1 >= 0), bit 1 = (k2 >= 0); q1 = floor(k / r), q2 = floor(k * (-v1) / r),
    k1 = k - q1*v0 - q2*u0, k2 = -q1*v1 - q2*u1.  The reference keeps only the low two 64-bit words of |k1|, |k2| (:133-136)."""
    q1 = k // GLV_DIVISOR
    q2 = (k * GLV_NEG_V1) // GLV_DIVISOR
    k1 = k - q1 - q2 * GLV_U0
    k2 = q1 * GLV_NEG_V1 - q2
    sign = (1 if k1 >= 0 else 0) | (2 if k2 >= 0 else 0)
    m = (1 << 128) - 1
    return abs(k1) & m, abs(k2) & m, sign


def glv_preprocess(points, scalars):
    """g1m_glv_preprocessEndomorphism (build_glv.js:178-263): N points/scalars -> 2N points, 2N scalars (each < 2^128)."""
    cv = BLS12_381
    out_p, out_s = [], []
    for P, k in zip(points, scalars):
        k1, k2, sign = glv_decompose(k)
        x, y = P
        out_p.append((x, y if sign & 1 else (-y) % cv.q))
        out_p.append((GLV_BETA * x % cv.q, y if sign & 2 else (-y) % cv.q))     # g1m_glv_endomorphism :150-174
        out_s += [k1, k2]
    return out_p, out_s


GLV_LAMBDA = BLS_LAMBDA      # phi(x, y) = (GLV_BETA * x, y) = GLV_LAMBDA * (x, y); signed halves satisfy k = k1 + k2 * GLV_LAMBDA (mod r)


# ---- G2: E'(Fq2), Fq2 = Fq[u]/(u^2 + 1) (build_f2m.js; f1m_neg is the non-residue multiplication on both curves).
# Elements are pairs (c0, c1) of plain integers; points are (x, y) pairs of such or None.  Used to pin the G2 path on small cases.
def f2_add(cv, a, b): return ((a[0] + b[0]) % cv.q, (a[1] + b[1]) % cv.q)
def f2_sub(cv, a, b): return ((a[0] - b[0]) % cv.q, (a[1] - b[1]) % cv.q)
def f2_mul(cv, a, b): return ((a[0] * b[0] - a[1] * b[1]) % cv.q, (a[0] * b[1] + a[1] * b[0]) % cv.q)        # f2m_mul, build_f2m.js:152-194
def f2_inv(cv, a):                                                                                            # f2m_inverse, build_f2m.js:402-440
    t = pow((a[0] * a[0] + a[1] * a[1]) % cv.q, -1, cv.q)
    return (a[0] * t % cv.q, (-a[1]) * t % cv.q)


def g2_add(cv, P, Q):
    if P is None: return Q
    if Q is None: return P
    if P[0] == Q[0]:
        if f2_add(cv, P[1], Q[1]) == (0, 0): return None
        lam = f2_mul(cv, f2_mul(cv, (3, 0), f2_mul(cv, P[0], P[0])), f2_inv(cv, f2_add(cv, P[1], P[1])))
    else:
        lam = f2_mul(cv, f2_sub(cv, Q[1], P[1]), f2_inv(cv, f2_sub(cv, Q[0], P[0])))
    x3 = f2_sub(cv, f2_sub(cv, f2_mul(cv, lam, lam), P[0]), Q[0])
    return (x3, f2_sub(cv, f2_mul(cv, lam, f2_sub(cv, P[0], x3)), P[1]))


def g2_mul(cv, k, P):
    acc = None
    while k:
        if k & 1: acc = g2_add(cv, acc, P)
        P = g2_add(cv, P, P); k >>= 1
    return acc


def g2_msm_naive(cv, points, scalars):
    acc = None
    for P, k in zip(points, scalars): acc = g2_add(cv, acc, g2_mul(cv, k, P))
    return acc


def g2_from_bytes(cv, b):
    """affine Montgomery bytes x0 || x1 || y0 || y1 -> point (or None for all-zero)"""
    n8 = cv.n8; Ri = pow(cv.R, -1, cv.q)
    v = [int.from_bytes(b[i * n8:(i + 1) * n8], "little") * Ri % cv.q for i in range(4)]
    return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))


def g2_canonical_bytes(cv, P):
    n8 = cv.n8
    if P is None: return bytes(4 * n8)
    return b"".join(c.to_bytes(n8, "little") for c in (P[0][0], P[0][1], P[1][0], P[1][1]))
