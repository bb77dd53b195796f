"""Host-side mirror of the reference's test rig for the MSM path.

The reference drives its WASM module through `wasmbuilder.buildProtoboard` (usage:
wasmcurves/test/batchAffine.js:13-17, benchmarks/multiexp.js:9-31): one linear memory, `pb.alloc`,
`pb.set`, `pb.get`, and every export as a method taking i32 "pointers".  This class offers the same
names and argument meaning for the exports on the MSM path, with the work done by the GPU engine
through the C ABI, so a parity test reads like the reference's own test:

    pb = Protoboard("bls12381")
    pBases = pb.alloc(n * 96); pScalars = pb.alloc(n * 32); pRes = pb.alloc(144)
    ... pb.set(...) / pb.f1m_toMontgomery(p, p) ...
    pb.g1m_multiexpAffine(pBases, pScalars, 32, n, pRes)
    pb.g1m_normalize(pRes, pRes); pb.f1m_fromMontgomery(pRes, pRes); pb.f1m_fromMontgomery(pRes + 48, pRes + 48)
    x, y = pb.get(pRes, 2, 48)

"Pointers" are byte offsets into a host bytearray (the linear memory); errors are raised as
B200MsmError where the reference would trap.
"""
import ctypes
from ._lib import BLS12_381_G1, BN254_G1, N8, B200MsmError
from .engine import Engine

_OPS = {"mul": 0, "add": 1, "sub": 2, "square": 3, "inverse": 4, "toMontgomery": 5, "fromMontgomery": 6, "neg": 7}


class Protoboard:
    def __init__(self, curve="bls12381", mem_bytes=1 << 26, engine=None):
        self.curve = BLS12_381_G1 if curve in ("bls12381", 0) else BN254_G1
        self.n8 = N8[self.curve]
        self.mem = bytearray(mem_bytes)
        self._cbuf = (ctypes.c_char * mem_bytes).from_buffer(self.mem)
        self._base = ctypes.addressof(self._cbuf)
        self._top = 8
        self.engine = engine or Engine()

    # ---- memory (pb.alloc / pb.set / pb.get; bump pointer like the data segment at address 0)
    def alloc(self, nbytes):
        p = (self._top + 7) & ~7
        if p + nbytes > len(self.mem): raise MemoryError("protoboard memory exhausted")
        self._top = p + nbytes
        return p

    def set(self, ptr, value, nbytes=4):
        self.mem[ptr:ptr + nbytes] = int(value).to_bytes(nbytes, "little")

    def get(self, ptr, count=1, nbytes=4):
        v = [int.from_bytes(self.mem[ptr + i * nbytes: ptr + (i + 1) * nbytes], "little") for i in range(count)]
        return v[0] if count == 1 else v

    def write(self, ptr, data): self.mem[ptr:ptr + len(data)] = data
    def read(self, ptr, n): return bytes(self.mem[ptr:ptr + n])
    def _p(self, ptr): return self._base + ptr

    # ---- MSM exports (wasmcurves/src/build_multiexp.js:251-371, :96-249; build_multiexp_opt.js:1987-2110)
    def g1m_multiexpAffine(self, pBases, pScalars, scalarSize, n, pr):
        self.engine.multiexp_affine(self.curve, self._p(pBases), self._p(pScalars), scalarSize, n, out=self._p(pr))

    g1m_multiexpAffine_wasmcurve = g1m_multiexpAffine      # the fork's name for the same export (build_curve_jacobian_a0.js:1429-1430)

    def g1m_multiexpAffine_chunk(self, pBases, pScalars, scalarSize, n, startBit, chunkSize, pr):
        out = self.engine.multiexp_affine_chunk(self.curve, self._p(pBases), self._p(pScalars), scalarSize, n, startBit, chunkSize)
        self.write(pr, out)

    g1m_multiexpAffine_wasmcurve_chunk = g1m_multiexpAffine_chunk

    def g1m_multiexp_multiExp(self, pPoints, pScalars, numPoints, pResult):
        """Manta entry point: affine points, 32-byte scalars (build_multiexp_opt.js:1987-1996, :2024)."""
        self.g1m_multiexpAffine(pPoints, pScalars, 32, numPoints, pResult)

    g1m_multiexpAffine_multiExp = g1m_multiexp_multiExp

    # ---- Jacobian bases (n8b = 3*n8, build_curve_jacobian_a0.js:1429) and the G2 exports (Fq2 elements of 2*n8 bytes)
    def g1m_multiexp(self, pBases, pScalars, scalarSize, n, pr):
        self.write(pr, self.engine.multiexp_jacobian(self.curve, self.read(pBases, n * 3 * self.n8), self.read(pScalars, n * scalarSize), scalarSize, n))

    def g1m_multiexp_chunk(self, pBases, pScalars, scalarSize, n, startBit, chunkSize, pr):
        self.write(pr, self.engine.multiexp_jacobian(self.curve, self.read(pBases, n * 3 * self.n8), self.read(pScalars, n * scalarSize), scalarSize, n, (startBit, chunkSize)))

    def g2m_multiexpAffine(self, pBases, pScalars, scalarSize, n, pr):
        self.engine.multiexp_affine(self.curve + 2, self._p(pBases), self._p(pScalars), scalarSize, n, out=self._p(pr))

    def g2m_multiexpAffine_chunk(self, pBases, pScalars, scalarSize, n, startBit, chunkSize, pr):
        self.write(pr, self.engine.multiexp_affine_chunk(self.curve + 2, self._p(pBases), self._p(pScalars), scalarSize, n, startBit, chunkSize))

    def g2m_multiexp(self, pBases, pScalars, scalarSize, n, pr):
        self.write(pr, self.engine.multiexp_jacobian(self.curve + 2, self.read(pBases, n * 6 * self.n8), self.read(pScalars, n * scalarSize), scalarSize, n))

    # ---- Fr transforms (wasmcurves/src/build_fft.js:178-245): in place on n Montgomery elements of 32 bytes
    def frm_fft(self, px, n): self._fft(px, n, False)
    def frm_ifft(self, px, n): self._fft(px, n, True)

    def _fft(self, px, n, inverse):
        lg = n.bit_length() - 1
        if n <= 0 or (1 << lg) != n: raise B200MsmError(-1, "frm_fft: n must be a power of two")
        self.engine.fr_fft(self.curve, self._p(px), lg, inverse=inverse, out=self._p(px))

    # ---- GLV pre-pass (wasmcurves/src/build_glv.js:53-146, 178-263; BLS12-381 only, like the reference)
    def g1m_glv_decomposeScalar(self, pScalar, pScalarRes):
        out, signs = self.engine.glv_decompose_scalars(self.curve, self.read(pScalar, 32), 1)
        self.write(pScalarRes, out); return signs[0]

    def g1m_glv_preprocessEndomorphism(self, pPoints, pScalars, numPoints, pPointsRes, pScalarsRes):
        pts, scs = self.engine.glv_preprocess(self.curve, self.read(pPoints, numPoints * 2 * self.n8), self.read(pScalars, numPoints * 32), numPoints)
        self.write(pPointsRes, pts); self.write(pScalarsRes, scs)

    # ---- batch conversions / codecs (wasmcurves/src/build_curve_jacobian_a0.js:1040-1328,1413-1418): (pIn, n, pOut)
    def _batch(self, op, pIn, n, pOut, in_sz, out_sz):
        self.write(pOut, self.engine.batch_convert(self.curve, op, self.read(pIn, n * in_sz), n))

    def g1m_batchLEMtoU(self, pIn, n, pOut): self._batch("LEMtoU", pIn, n, pOut, 2 * self.n8, 2 * self.n8)
    def g1m_batchUtoLEM(self, pIn, n, pOut): self._batch("UtoLEM", pIn, n, pOut, 2 * self.n8, 2 * self.n8)
    def g1m_batchLEMtoC(self, pIn, n, pOut): self._batch("LEMtoC", pIn, n, pOut, 2 * self.n8, self.n8)
    def g1m_batchCtoLEM(self, pIn, n, pOut): self._batch("CtoLEM", pIn, n, pOut, self.n8, 2 * self.n8)
    def g1m_batchToAffine(self, pIn, n, pOut): self._batch("toAffine", pIn, n, pOut, 3 * self.n8, 2 * self.n8)
    def g1m_batchToJacobian(self, pIn, n, pOut): self._batch("toJacobian", pIn, n, pOut, 2 * self.n8, 3 * self.n8)

    # ---- the helpers every reference test uses around the MSM call
    def _fq(self, op, pa, pb_, pr):
        a = self.read(pa, self.n8); b = self.read(pb_, self.n8) if pb_ is not None else None
        self.write(pr, self.engine.fq_op(self.curve, _OPS[op], a, b))

    def f1m_toMontgomery(self, pa, pr): self._fq("toMontgomery", pa, None, pr)
    def f1m_fromMontgomery(self, pa, pr): self._fq("fromMontgomery", pa, None, pr)
    def f1m_mul(self, pa, pb_, pr): self._fq("mul", pa, pb_, pr)
    def f1m_add(self, pa, pb_, pr): self._fq("add", pa, pb_, pr)
    def f1m_sub(self, pa, pb_, pr): self._fq("sub", pa, pb_, pr)
    def f1m_square(self, pa, pr): self._fq("square", pa, None, pr)
    def f1m_inverse(self, pa, pr): self._fq("inverse", pa, None, pr)
    def f1m_neg(self, pa, pr): self._fq("neg", pa, None, pr)

    def g1m_isZero(self, p):
        return int(self.read(p + 2 * self.n8, self.n8) == bytes(self.n8))

    def g1m_normalize(self, p, pr):
        """Jacobian -> Jacobian with z = 1 (Montgomery), infinity -> g1m_zero (build_curve_jacobian_a0.js:940-973)."""
        n8 = self.n8
        if self.g1m_isZero(p):
            one = self.engine.fq_op(self.curve, _OPS["toMontgomery"], (1).to_bytes(n8, "little"))
            self.write(pr, bytes(n8) + one + bytes(n8)); return
        xy = self.engine.normalize(self.curve, self.read(p, 3 * n8))          # canonical, non-Montgomery
        xm = self.engine.fq_op(self.curve, _OPS["toMontgomery"], xy[:n8]); ym = self.engine.fq_op(self.curve, _OPS["toMontgomery"], xy[n8:])
        one = self.engine.fq_op(self.curve, _OPS["toMontgomery"], (1).to_bytes(n8, "little"))
        self.write(pr, xm + ym + one)

    def g1m_add(self, p1, p2, pr):
        n8 = self.n8
        self.write(pr, self.engine.sum_points(self.curve, self.read(p1, 3 * n8) + self.read(p2, 3 * n8), 2))
