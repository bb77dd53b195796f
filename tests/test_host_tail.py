"""CPU tests of the engine's host-side arithmetic (zprize-wasm-msm_b200/csrc/host_ec.h): the serial window combination that
finishes every MSM (accumulateAcrossChunks / the multiexp Horner loop of the reference, build_multiexp_opt.js:1710-1746,
build_multiexp.js:319-369) and its sub-slot form for the window-table path, over Fq (G1) and Fq2 (G2), both curves.
The header is compiled with g++ into a small harness; expectations come from big-integer affine arithmetic (oracle/pyref.py)."""
import ctypes, os, random, subprocess
import pytest
import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_host_ec_harness.so")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "host_ec_harness.cpp"); hdr = os.path.join(ROOT, "zprize-wasm-msm_b200", "csrc", "host_ec.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.dirname(hdr), "-o", SO, src])
    lib = ctypes.CDLL(SO)
    lib.host_tail.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_uint32] * 6 + [ctypes.c_void_p]
    lib.host_mul.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lib


class Grp:
    """group arithmetic + byte formats for (curve, ext): points are pyref affine tuples (G1) or pairs of Fq2 pairs (G2)"""

    def __init__(self, cname, ext):
        self.cv = pyref.CURVES[cname]; self.ext = ext; self.n8 = self.cv.n8; self.words = self.n8 // 8
        self.qb = self.cv.q.to_bytes(self.n8, "little"); self.oneb = (self.cv.R % self.cv.q).to_bytes(self.n8, "little")
        if ext == 1: self.G = self.cv.G
        else:
            g2 = {"bls12381": ((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
                                3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
                               (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
                                927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582)),
                  "bn128": ((10857046999023057135944570762232829481370756359578518086990519993285655852781, 11559732032986387107991004021392285783925812861821192530917403151452391805634),
                            (8495653923123431417604973247489272438418190587263600148770280649306958101930, 4082367875863433681332203403145435568316851327593401208105741076214120093531))}
            self.G = g2[cname]                                     # build_bls12381.js:127-138, build_bn128.js:122-133

    def add(self, P, Q): return pyref.add(self.cv, P, Q) if self.ext == 1 else pyref.g2_add(self.cv, P, Q)
    def mul(self, k, P): return pyref.mul(self.cv, k, P) if self.ext == 1 else pyref.g2_mul(self.cv, k, P)

    def fe(self, c):
        """field element (int for Fq, pair for Fq2) -> Montgomery bytes"""
        cs = (c,) if self.ext == 1 else c
        return b"".join((x * self.cv.R % self.cv.q).to_bytes(self.n8, "little") for x in cs)

    def xyzz(self, P):
        """affine point -> XYZZ Montgomery bytes with zz = zzz = 1; None -> zz = 0"""
        one = 1 if self.ext == 1 else (1, 0); zero = 0 if self.ext == 1 else (0, 0)
        if P is None: return self.fe(zero) + self.fe(one) + self.fe(zero) + self.fe(zero)
        return self.fe(P[0]) + self.fe(P[1]) + self.fe(one) + self.fe(one)

    def from_jac(self, b):
        """Jacobian Montgomery bytes -> affine point or None"""
        q = self.cv.q; Ri = pow(self.cv.R, -1, q); e = self.ext * self.n8
        def rd(o): v = [int.from_bytes(b[o + i * self.n8: o + (i + 1) * self.n8], "little") * Ri % q for i in range(self.ext)]; return v[0] if self.ext == 1 else tuple(v)
        X, Y, Z = rd(0), rd(e), rd(2 * e)
        if self.ext == 1:
            if Z == 0: return None
            zi = pow(Z, -1, q); return (X * zi * zi % q, Y * zi * zi * zi % q)
        if Z == (0, 0): return None
        cv = self.cv; zi = pyref.f2_inv(cv, Z); zi2 = pyref.f2_mul(cv, zi, zi)
        return (pyref.f2_mul(cv, X, zi2), pyref.f2_mul(cv, Y, pyref.f2_mul(cv, zi2, zi)))


CASES = [("bls12381", 1), ("bn128", 1), ("bls12381", 2), ("bn128", 2)]


@pytest.mark.parametrize("cname,ext", CASES)
def test_host_field_mul_and_square(harness, cname, ext):
    g = Grp(cname, ext); cv = g.cv; rnd = random.Random(ext * 7 + len(cname))
    def rand(): return rnd.randrange(cv.q) if ext == 1 else (rnd.randrange(cv.q), rnd.randrange(cv.q))
    def fmul(a, b): return a * b % cv.q if ext == 1 else pyref.f2_mul(cv, a, b)
    edge = [0, 1, cv.q - 1] if ext == 1 else [(0, 0), (1, 0), (0, 1), (cv.q - 1, cv.q - 1)]
    vals = edge + [rand() for _ in range(40)]
    for a in vals:
        for b in (vals[3], vals[-1], a):
            out = ctypes.create_string_buffer(ext * g.n8)
            assert harness.host_mul(ext, g.words, g.qb, g.oneb, g.fe(a), g.fe(b), out, 0) == 0
            assert out.raw == g.fe(fmul(a, b))
        out = ctypes.create_string_buffer(ext * g.n8)
        assert harness.host_mul(ext, g.words, g.qb, g.oneb, g.fe(a), g.fe(a), out, 1) == 0
        assert out.raw == g.fe(fmul(a, a))


@pytest.mark.parametrize("cname,ext", CASES)
@pytest.mark.parametrize("Wd,c0,rem", [(5, 4, 0), (6, 3, 2), (1, 7, 0), (16, 16, 0)])
def test_window_combiner(harness, cname, ext, Wd, c0, rem):
    """result = sum_w 2^(off_w) * (T_w[0] + sum_j 2^j T_w[2^j]) (+ the extra slot of the last window when all widths are equal),
    fed whole and as two groups from the top down; includes infinity entries."""
    g = Grp(cname, ext); rnd = random.Random(Wd * 100 + c0 + ext)
    c = c0 + (1 if rem else 0); logB = c - 1; W = Wd + (1 if rem == 0 else 0); per = logB + 1
    def off(w): return w * c0 + min(w, rem)
    small = [g.mul(k, g.G) for k in range(1, 12)]
    slots = [[(None if rnd.random() < 0.2 else small[rnd.randrange(len(small))]) for _ in range(per)] for _ in range(W)]
    exp = None
    for w in range(Wd):
        v = slots[w][0]
        for j in range(logB): v = g.add(v, g.mul(1 << j, slots[w][1 + j])) if slots[w][1 + j] is not None else v
        exp = g.add(exp, g.mul(1 << off(w), v)) if v is not None else exp
    if W > Wd:   # extra slot: buckets B+1 .. 2B of the last window: (2^logB + 1) E[0] + sum_j 2^j E[2^j]
        E = slots[Wd]; v = g.mul((1 << logB) + 1, E[0]) if E[0] is not None else None
        for j in range(logB): v = g.add(v, g.mul(1 << j, E[1 + j])) if E[1 + j] is not None else v
        exp = g.add(exp, g.mul(1 << off(Wd - 1), v)) if v is not None else exp
    folded = b"".join(g.xyzz(p) for s in slots for p in s)
    for split in (0, max(1, Wd // 2)):
        out = ctypes.create_string_buffer(3 * ext * g.n8)
        assert harness.host_tail(ext, g.words, g.qb, g.oneb, 0, folded, W, Wd, c0, rem, logB, split if split < Wd else 0, out) == 0
        assert g.from_jac(out.raw) == exp, split


@pytest.mark.parametrize("cname,ext", CASES)
@pytest.mark.parametrize("S,logBs", [(1, 5), (4, 6), (8, 3), (2, 1)])
def test_subslot_combiner(harness, cname, ext, S, logBs):
    """window-table form: sum_b (b+1) T[b] over S sub-slots of 2^logBs buckets from their folded entries:
    V_s = F_s[0] + sum_j 2^j F_s[2^j] + s * 2^logBs * F_s[0]"""
    g = Grp(cname, ext); rnd = random.Random(S * 10 + logBs + ext); per = logBs + 1
    small = [g.mul(k, g.G) for k in range(1, 12)]
    slots = [[(None if rnd.random() < 0.25 else small[rnd.randrange(len(small))]) for _ in range(per)] for _ in range(S)]
    exp = None
    for s in range(S):
        F = slots[s]; v = g.mul(1 + (s << logBs), F[0]) if F[0] is not None else None
        for j in range(logBs): v = g.add(v, g.mul(1 << j, F[1 + j])) if F[1 + j] is not None else v
        exp = g.add(exp, v)
    folded = b"".join(g.xyzz(p) for s in slots for p in s)
    for split in (0, S // 2):
        out = ctypes.create_string_buffer(3 * ext * g.n8)
        assert harness.host_tail(ext, g.words, g.qb, g.oneb, 1, folded, S, logBs, 0, 0, 0, split, out) == 0
        assert g.from_jac(out.raw) == exp, split
