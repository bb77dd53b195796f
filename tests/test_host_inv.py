"""CPU test of zprize-wasm-msm_b200/csrc/bingcd.h -- the modular inversion at the root of every batch inversion
(f1m_inverse, wasmcurves/src/build_f1m.js:1112-1122 / build_int.js:922-1064; used by f1m_batchInverse, build_batchinverse.js:90).
The header is plain C++: g++ compiles here exactly what nvcc compiles into the engine (fp.cuh: fe_inv_fast_p)."""
import ctypes, os, random, subprocess
import pytest

HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_host_inv_harness.so")
FIELDS = {0: "BLS12-381 Fq", 1: "BN254 Fq", 2: "BLS12-381 Fr", 3: "BN254 Fr"}


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "host_inv_harness.cpp"); inc = os.path.join(ROOT, "zprize-wasm-msm_b200", "csrc")
    deps = [src, os.path.join(inc, "bingcd.h"), os.path.join(inc, "field_params.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", inc, "-o", SO, src])
    return ctypes.CDLL(SO)


def _field(h, field):
    n = ctypes.c_uint32(); qb = ctypes.create_string_buffer(64)
    assert h.host_field_info(field, ctypes.byref(n), qb) == 0
    return n.value, int.from_bytes(qb.raw[: 4 * n.value], "little")


def _inv(h, field, N, vals):
    inb = b"".join(v.to_bytes(4 * N, "little") for v in vals); out = ctypes.create_string_buffer(max(1, len(inb)))
    assert h.host_inverse(field, inb, out, len(vals)) == 0
    return [int.from_bytes(out.raw[i * 4 * N:(i + 1) * 4 * N], "little") for i in range(len(vals))]


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_field_constants_are_the_reference_moduli(harness, field):
    import pyref
    N, q = _field(harness, field)
    exp = {0: pyref.BLS12_381.q, 1: pyref.BN254.q, 2: pyref.BLS12_381.r, 3: pyref.BN254.r}[field]
    assert q == exp and N == (12 if field == 0 else 8)


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_bingcd_inverse_edges_and_random(harness, field):
    N, q = _field(harness, field); rnd = random.Random(100 + field); L = q.bit_length()
    vals = [0, 1, 2, 3, q - 1, q - 2, (q - 1) // 2, (q + 1) // 2, 1 << 31, 1 << 32, 1 << 33, 1 << 63, 1 << 64, 1 << 65, 1 << (L - 1), (1 << (L - 1)) - 1]
    vals += [1 << k for k in range(0, L - 1)] + [(1 << k) - 1 for k in range(1, L - 1)] + [q - (1 << k) for k in range(0, L - 1, 3)]
    vals += [rnd.randrange(q) for _ in range(20000)] + [rnd.randrange(1 << k) for k in range(1, L - 1) for _ in range(4)]
    vals = [v % q for v in vals]
    got = _inv(harness, field, N, vals)
    for v, g in zip(vals, got):
        assert g == (pow(v, -1, q) if v else 0), hex(v)


@pytest.mark.parametrize("field", [0, 1])
def test_bingcd_inverse_is_an_involution_on_montgomery_shaped_values(harness, field):
    """the engine feeds it a*R mod q; inv(inv(x)) == x and x * inv(x) == 1 on such values"""
    N, q = _field(harness, field); rnd = random.Random(7); R = 1 << (32 * N)
    xs = [rnd.randrange(1, q) * R % q for _ in range(3000)]
    inv = _inv(harness, field, N, xs)
    assert all(x * y % q == 1 for x, y in zip(xs, inv))
    assert _inv(harness, field, N, inv) == xs
