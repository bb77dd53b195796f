"""refwasm.py -- TEST INFRASTRUCTURE (oracle), not product code.

ctypes rig around oracle/_ref/libref_<curve>.so, i.e. the reference's own shipped
WASM module (wasmcurves/build/{bls12381,bn128}.wasm) translated to C by
oracle/wasm2c.py and compiled natively.  It mirrors the parts of
wasmbuilder.buildProtoboard that the reference's tests use
(test/batchAffine.js:13-17, benchmarks/multiexp.js:9-31): pb.alloc / pb.set /
pb.get / every export as a method, all data in one linear memory.

Nothing here is importable from the product path.
"""
import ctypes, os, re

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

# constant pointers from build/<curve>_wasm.js (tools/buildwasm_bls12381.js:14-31); read at
# build time by oracle/build_ref.py and stored beside the .so so /root/reference is not needed at run time.


def available(curve="bls12381"):
    return os.path.exists(os.path.join(_REF, "libref_%s.so" % curve))


class RefModule:
    """One instance of the reference module (single-threaded, like a WASM instance)."""

    def __init__(self, curve="bls12381", pages=4096):
        path = os.path.join(_REF, "libref_%s.so" % curve)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `python oracle/build_ref.py` where /root/reference exists)")
        # private copy of the symbols: several instances/curves may coexist
        self.lib = ctypes.CDLL(path, mode=os.RTLD_LOCAL)
        self.curve = curve
        self.n8 = 48 if curve == "bls12381" else 32
        self.lib.wasm_mem.restype = ctypes.c_void_p
        self.lib.wasm_mem_bytes.restype = ctypes.c_uint32
        assert self.lib.wasm_init(ctypes.c_uint32(pages)) == 0
        self.base = self.lib.wasm_mem()
        self.nbytes = self.lib.wasm_mem_bytes()
        self.consts = {}
        cpath = os.path.join(_REF, "%s_consts.txt" % curve)
        if os.path.exists(cpath):
            for line in open(cpath):
                k, v = line.split()
                self.consts[k] = int(v)

    # ---- protoboard-like memory API
    def _buf(self, ptr, n):
        assert 0 <= ptr and ptr + n <= self.nbytes, "out of linear memory"
        return (ctypes.c_uint8 * n).from_address(self.base + ptr)

    def alloc(self, nbytes):
        """pb.alloc: bump pointer stored at address 0, 8-byte aligned (wasmbuilder protoboard semantics)."""
        p = int.from_bytes(bytes(self._buf(0, 4)), "little")
        while p & 7: p += 1
        new = p + nbytes
        assert new <= self.nbytes, "reference instance out of memory"
        self._buf(0, 4)[:] = new.to_bytes(4, "little")
        return p

    def heap_mark(self):
        return int.from_bytes(bytes(self._buf(0, 4)), "little")

    def heap_release(self, mark):
        self._buf(0, 4)[:] = mark.to_bytes(4, "little")

    def write(self, ptr, data):
        ctypes.memmove(self.base + ptr, bytes(data), len(data))

    def read(self, ptr, n):
        return bytes(self._buf(ptr, n))

    def set(self, ptr, value, nbytes=4):
        self.write(ptr, int(value).to_bytes(nbytes, "little"))

    def get(self, ptr, count=1, nbytes=4):
        v = [int.from_bytes(self.read(ptr + i * nbytes, nbytes), "little") for i in range(count)]
        return v[0] if count == 1 else v

    def __getattr__(self, name):
        if name.startswith("_"): raise AttributeError(name)
        fn = getattr(self.lib, name)
        fn.restype = ctypes.c_uint32

        def call(*args):
            return fn(*[ctypes.c_uint32(a) for a in args])
        return call

    # ---- helpers used by tests / bench (compose only reference exports)
    def msm_affine(self, bases_mont: bytes, scalars: bytes, scalar_size: int, n: int):
        """g1m_multiexpAffine -> canonical (x, y) ints or None; exactly test/batchAffine.js:1222-1254."""
        n8 = self.n8
        mark = self.heap_mark()
        pB = self.alloc(max(len(bases_mont), 8)); pS = self.alloc(len(scalars) + 8); pR = self.alloc(3 * n8)
        self.write(pB, bases_mont); self.write(pS, scalars)
        self.g1m_multiexpAffine(pB, pS, scalar_size, n, pR)
        out = self.normalize_read(pR)
        self.heap_release(mark)
        return out

    def msm_affine_raw(self, bases_mont: bytes, scalars: bytes, scalar_size: int, n: int) -> bytes:
        """g1m_multiexpAffine -> raw Jacobian-Montgomery bytes (3*n8)."""
        n8 = self.n8
        mark = self.heap_mark()
        pB = self.alloc(max(len(bases_mont), 8)); pS = self.alloc(len(scalars) + 8); pR = self.alloc(3 * n8)
        self.write(pB, bases_mont); self.write(pS, scalars)
        self.g1m_multiexpAffine(pB, pS, scalar_size, n, pR)
        out = self.read(pR, 3 * n8)
        self.heap_release(mark)
        return out

    def msm_chunk(self, bases_mont, scalars, scalar_size, n, start_bit, chunk_bits):
        n8 = self.n8
        mark = self.heap_mark()
        pB = self.alloc(max(len(bases_mont), 8)); pS = self.alloc(len(scalars) + 8); pR = self.alloc(3 * n8)
        self.write(pB, bases_mont); self.write(pS, scalars)
        self.g1m_multiexpAffine_chunk(pB, pS, scalar_size, n, start_bit, chunk_bits, pR)
        out = self.normalize_read(pR)
        self.heap_release(mark)
        return out

    def normalize_read(self, pR):
        n8 = self.n8
        if self.g1m_isZero(pR): return None
        self.g1m_normalize(pR, pR)
        self.f1m_fromMontgomery(pR, pR); self.f1m_fromMontgomery(pR + n8, pR + n8)
        x, y = self.get(pR, 2, n8)
        return (x, y)

    def normalize_bytes(self, jac: bytes):
        mark = self.heap_mark()
        pR = self.alloc(3 * self.n8); self.write(pR, jac)
        out = self.normalize_read(pR)
        self.heap_release(mark)
        return out


# ---- G2 (g2m_* exports over f2m, build_bls12381.js:48-53 / build_bn128.js:44-49): same rig, elements are Fq2 = c0 || c1
class RefG2:
    """helpers over one RefModule for the G2 exports; all data in the reference's own byte formats"""

    def __init__(self, mod: RefModule):
        self.m = mod; self.n8 = mod.n8; self.e8 = 2 * mod.n8          # bytes per Fq / per Fq2 element

    def generator_affine(self) -> bytes:
        return self.m.read(self.m.consts["pG2gen"], 2 * self.e8)

    def times_scalar_affine(self, base_affine: bytes, scalar: bytes) -> bytes:
        """g2m_timesScalarAffine + g2m_toAffine -> affine Montgomery bytes (2 * e8)"""
        m = self.m; mark = m.heap_mark()
        pB = m.alloc(2 * self.e8); pS = m.alloc(len(scalar) + 8); pR = m.alloc(3 * self.e8)
        m.write(pB, base_affine); m.write(pS, scalar)
        m.g2m_timesScalarAffine(pB, pS, len(scalar), pR)
        m.g2m_toAffine(pR, pR)
        out = m.read(pR, 2 * self.e8); m.heap_release(mark); return out

    def canonical(self, pR) -> bytes:
        """g2m_normalize + f2m_fromMontgomery on x and y -> x0 || x1 || y0 || y1 as plain LE integers; infinity -> zeros"""
        m = self.m
        if m.g2m_isZero(pR): return bytes(2 * self.e8)
        m.g2m_normalize(pR, pR)
        m.f2m_fromMontgomery(pR, pR); m.f2m_fromMontgomery(pR + self.e8, pR + self.e8)
        return m.read(pR, 2 * self.e8)

    def msm_affine(self, bases: bytes, scalars: bytes, scalar_size: int, n: int) -> bytes:
        m = self.m; mark = m.heap_mark()
        pB = m.alloc(max(len(bases), 8)); pS = m.alloc(len(scalars) + 8); pR = m.alloc(3 * self.e8)
        m.write(pB, bases); m.write(pS, scalars)
        m.g2m_multiexpAffine(pB, pS, scalar_size, n, pR)
        out = self.canonical(pR); m.heap_release(mark); return out

    def msm_chunk(self, bases, scalars, scalar_size, n, start_bit, chunk_bits) -> bytes:
        m = self.m; mark = m.heap_mark()
        pB = m.alloc(max(len(bases), 8)); pS = m.alloc(len(scalars) + 8); pR = m.alloc(3 * self.e8)
        m.write(pB, bases); m.write(pS, scalars)
        m.g2m_multiexpAffine_chunk(pB, pS, scalar_size, n, start_bit, chunk_bits, pR)
        out = self.canonical(pR); m.heap_release(mark); return out

    def canonical_of(self, jac: bytes) -> bytes:
        m = self.m; mark = m.heap_mark()
        pR = m.alloc(3 * self.e8); m.write(pR, jac)
        out = self.canonical(pR); m.heap_release(mark); return out

    def f2m(self, fn: str, a: bytes, b: bytes = None) -> bytes:
        """one f2m_<fn> call on Montgomery Fq2 operands"""
        m = self.m; mark = m.heap_mark()
        pA = m.alloc(self.e8); pB = m.alloc(self.e8); pR = m.alloc(self.e8)
        m.write(pA, a)
        if b is not None: m.write(pB, b); getattr(m, "f2m_" + fn)(pA, pB, pR)
        else: getattr(m, "f2m_" + fn)(pA, pR)
        out = m.read(pR, self.e8); m.heap_release(mark); return out


def msm_jacobian(mod: RefModule, group: str, bases_jac: bytes, scalars: bytes, scalar_size: int, n: int, chunk=None) -> bytes:
    """<group>_multiexp / <group>_multiexp_chunk of the reference module over JACOBIAN bases -> raw Jacobian result bytes.
    group = "g1m" (elements of n8 bytes) or "g2m" (2*n8)."""
    e8 = mod.n8 * (2 if group == "g2m" else 1)
    mark = mod.heap_mark()
    pB = mod.alloc(max(len(bases_jac), 8)); pS = mod.alloc(len(scalars) + 8); pR = mod.alloc(3 * e8)
    mod.write(pB, bases_jac); mod.write(pS, scalars)
    if chunk is None: getattr(mod, group + "_multiexp")(pB, pS, scalar_size, n, pR)
    else: getattr(mod, group + "_multiexp_chunk")(pB, pS, scalar_size, n, chunk[0], chunk[1], pR)
    out = mod.read(pR, 3 * e8); mod.heap_release(mark); return out
