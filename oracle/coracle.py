"""coracle.py -- TEST INFRASTRUCTURE (oracle), not product code.

ctypes binding of oracle/liboracle.so (oracle/msm_oracle.c, the plain-C restatement of the
reference's upstream MSM path).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this.
"""
import ctypes, os, subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None
N8 = {0: 48, 1: 32}


def build(force=False):
    src = os.path.join(_HERE, "msm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO): build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_point_scalar.restype = ctypes.c_uint64
        _lib.oracle_point_scalar.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    return _lib


def _out(n): return ctypes.create_string_buffer(n)


def multiexp_affine(curve, bases, scalars, scalar_size, n):
    o = _out(3 * N8[curve])
    lib().oracle_multiexp_affine(curve, bases, scalars, ctypes.c_uint32(scalar_size), ctypes.c_uint64(n), o)
    return o.raw


def multiexp_affine_chunk(curve, bases, scalars, scalar_size, n, start_bit, chunk_bits):
    o = _out(3 * N8[curve])
    lib().oracle_multiexp_affine_chunk(curve, bases, scalars, ctypes.c_uint32(scalar_size), ctypes.c_uint64(n),
                                       ctypes.c_uint32(start_bit), ctypes.c_uint32(chunk_bits), o)
    return o.raw


def normalize(curve, jac):
    """Jacobian-Montgomery bytes -> canonical x||y bytes (plain LE ints; infinity = zeros)."""
    o = _out(2 * N8[curve]); lib().oracle_normalize(curve, bytes(jac), o); return o.raw


def add(curve, a, b):
    o = _out(3 * N8[curve]); lib().oracle_add(curve, bytes(a), bytes(b), o); return o.raw


def times_scalar_affine(curve, base_xy, k_bytes):
    o = _out(3 * N8[curve]); lib().oracle_times_scalar_affine(curve, bytes(base_xy), bytes(k_bytes), ctypes.c_uint32(len(k_bytes)), o)
    return o.raw


def generate_bases(curve, gen_xy, seed, first, n):
    o = _out(max(1, 2 * N8[curve] * n))
    lib().oracle_generate_bases(curve, bytes(gen_xy), ctypes.c_uint64(seed), ctypes.c_uint64(first), ctypes.c_uint64(n), o)
    return o.raw[:2 * N8[curve] * n]


def point_scalar(seed, i): return lib().oracle_point_scalar(seed, i)


def _fe2(name, curve, a, b):
    n = len(a) // N8[curve]; o = _out(max(1, len(a)))
    getattr(lib(), name)(curve, bytes(a), bytes(b), o, ctypes.c_uint64(n)); return o.raw[:len(a)]


def _fe1(name, curve, a):
    n = len(a) // N8[curve]; o = _out(max(1, len(a)))
    getattr(lib(), name)(curve, bytes(a), o, ctypes.c_uint64(n)); return o.raw[:len(a)]


def fe_mul(curve, a, b): return _fe2("oracle_fe_mul", curve, a, b)
def fe_add(curve, a, b): return _fe2("oracle_fe_add", curve, a, b)
def fe_sub(curve, a, b): return _fe2("oracle_fe_sub", curve, a, b)
def fe_inv(curve, a): return _fe1("oracle_fe_inv", curve, a)
def fe_to_mont(curve, a): return _fe1("oracle_fe_to_mont", curve, a)
def fe_from_mont(curve, a): return _fe1("oracle_fe_from_mont", curve, a)
