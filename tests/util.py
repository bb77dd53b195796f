"""Shared helpers for the parity tests: seeded synthetic inputs and oracle shortcuts."""
import random
import pyref, coracle

SEED0 = 0xB2000000


def curve(cname): return pyref.CURVES[cname]


def gen_bytes(cv): return pyref.affine_to_bytes(cv, cv.G)


def make_bases(cv, n, seed=SEED0):
    """P_i = k_i * G, k_i = splitmix64(seed + i): same stream as the engine's b200msm_g1_generate_bases."""
    return coracle.generate_bases(cv.cid, gen_bytes(cv), seed, 0, n)


def make_scalars(n, seed, kind="u256", r=None):
    rnd = random.Random(seed)
    if kind == "u256": vals = [rnd.getrandbits(256) for _ in range(n)]
    elif kind == "modr": vals = [rnd.randrange(r) for _ in range(n)]
    elif kind == "small": vals = [rnd.randrange(4) for _ in range(n)]
    elif kind == "equal": v = rnd.getrandbits(256); vals = [v] * n
    else: raise ValueError(kind)
    return b"".join(v.to_bytes(32, "little") for v in vals)


def oracle_msm(cv, bases, scalars, scalar_size, n):
    """canonical x||y bytes from the C oracle"""
    return coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, bases, scalars, scalar_size, n))
