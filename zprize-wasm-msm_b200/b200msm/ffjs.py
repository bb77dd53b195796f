"""ffjavascript-style group surface over the engine ("next").

snarkjs-class callers do not call the WASM exports directly; they go through ffjavascript's curve objects
(`curve.G1.multiExpAffine(buffBases, buffScalars)`, `G1.batchLEMtoU(buff)`, ...), which slice the points, run
`g1m_multiexpAffine_chunk` per window in a worker pool and Horner-combine (RECOLLECTION of the un-vendored
ffjavascript@0.2.56 `engine_multiexp.js` / `engine_batchconvert.js`; the WASM side of those calls is
wasmcurves/src/build_multiexp.js:96-249 and build_curve_jacobian_a0.js:1040-1328,1413-1418).
This class offers the same method names, argument meaning and error behaviour over byte buffers; the slicing and the
worker pool disappear because one engine call does the whole MSM on the GPU.
"""
from ._lib import BLS12_381_G1, BN254_G1, BLS12_381_G2, BN254_G2, N8


class G1:
    def __init__(self, engine, curve):
        self.engine = engine
        self.curve = BLS12_381_G1 if curve in ("bls12381", BLS12_381_G1) else BN254_G1
        self.n8 = N8[self.curve]

    # ---- engine_multiexp.js
    @staticmethod
    def _split(buffBases, buffScalars, sGIn):
        nPoints = len(buffBases) // sGIn
        if nPoints * sGIn != len(buffBases): raise ValueError("Base size does not match")
        if nPoints == 0: return 0, 0
        sScalar = len(buffScalars) // nPoints
        if sScalar * nPoints != len(buffScalars): raise ValueError("Scalar size does not match")
        return nPoints, sScalar

    def multiExpAffine(self, buffBases, buffScalars):
        """sum_i scalar_i * base_i; bases affine LEM (2*n8 bytes each), scalars little-endian, all the same size.
        Returns the point in Jacobian LEM form (3*n8 bytes), like ffjavascript's G1.multiExpAffine."""
        nPoints, sScalar = self._split(buffBases, buffScalars, 2 * self.n8)
        if nPoints == 0: return self.zero()
        return self.engine.multiexp_affine(self.curve, buffBases, buffScalars, sScalar, nPoints)

    def multiExp(self, buffBases, buffScalars):
        """Same with Jacobian bases (3*n8 bytes each): g1m_multiexp (converted to affine on the GPU, then the same pipeline)."""
        nPoints, sScalar = self._split(buffBases, buffScalars, 3 * self.n8)
        if nPoints == 0: return self.zero()
        return self.engine.multiexp_jacobian(self.curve, buffBases, buffScalars, sScalar, nPoints)

    # ---- engine_batchconvert.js
    def _conv(self, op, buff, in_sz):
        n = len(buff) // in_sz
        if n * in_sz != len(buff): raise ValueError("Invalid buffer size")
        return self.engine.batch_convert(self.curve, op, buff, n)

    def batchLEMtoU(self, buff): return self._conv("LEMtoU", buff, 2 * self.n8)
    def batchUtoLEM(self, buff): return self._conv("UtoLEM", buff, 2 * self.n8)
    def batchLEMtoC(self, buff): return self._conv("LEMtoC", buff, 2 * self.n8)
    def batchCtoLEM(self, buff): return self._conv("CtoLEM", buff, self.n8)
    def batchToAffine(self, buff): return self._conv("toAffine", buff, 3 * self.n8)
    def batchToJacobian(self, buff): return self._conv("toJacobian", buff, 2 * self.n8)

    # ---- small helpers of the curve object
    def zero(self):
        """g1m_zero: (0, R mod q, 0)"""
        from . import constants
        _, _, one, _, _ = constants(self.curve)
        return bytes(self.n8) + one.to_bytes(self.n8, "little") + bytes(self.n8)

    def toAffine(self, p): return self.engine.batch_convert(self.curve, "toAffine", p, 1)
    def add(self, a, b): return self.engine.sum_points(self.curve, bytes(a) + bytes(b), 2)
    def eq(self, a, b): return self.engine.normalize(self.curve, a) == self.engine.normalize(self.curve, b)
    def isZero(self, p): return bytes(p[2 * self.n8:3 * self.n8]) == bytes(self.n8)


class G2(G1):
    """curve.G2 of ffjavascript: the same multiexp surface over the G2 exports (g2m_multiexpAffine / g2m_multiexp); elements are
    Fq2 = c0 || c1, so n8 here is 96 / 64 bytes.  The wire codecs (batchLEMtoU ...) are the g2m_batch* exports: the same kernels over Fq2."""

    def __init__(self, engine, curve):
        self.engine = engine
        self.curve = BLS12_381_G2 if curve in ("bls12381", BLS12_381_G2) else BN254_G2
        self.n8 = N8[self.curve]

    def zero(self):
        """g2m_zero: (0, (R mod q) + 0u, 0)"""
        from . import constants
        _, _, one, _, _ = constants(self.curve - 2)
        h = self.n8 // 2
        return bytes(self.n8) + one.to_bytes(h, "little") + bytes(h) + bytes(self.n8)

