// codecs.cuh -- the data formats on either side of the MSM path ("next").
//
// GPU counterparts of the point codecs and batch conversions that an ffjavascript / snarkjs-class host feeds the MSM from
// (wasmcurves/src/build_curve_jacobian_a0.js):
//   g1m_batchLEMtoU :1209-1236,1413   affine LE-Montgomery  -> uncompressed big-endian plain (x||y), infinity = 0x40 flag
//   g1m_batchUtoLEM :1238-1263,1415   inverse
//   g1m_batchLEMtoC :1166-1207,1414   affine LE-Montgomery  -> compressed big-endian x, 0x80 = "y is the greater root", 0x40 = infinity
//   g1m_batchCtoLEM :1265-1328,1416   inverse (y = sqrt(x^3 + b), f1m_sqrt; both fields have q = 3 mod 4)
//   g1m_batchToAffine :1040-1125      Jacobian -> affine (shared inversion), infinity -> (0,0)
//   g1m_batchToJacobian :1418         affine -> Jacobian (z = 1), infinity -> g1m_zero
// The same kernels serve G2 (g2m_batch*, wired for every prefix at :1413-1418): elements are Fq2 = c0 || c1, the byte reversal covers the whole
// 2*n8 element (big-endian c1 first), the sign is f2m_sign (build_f2m.js:411-430: the sign of c1, of c0 when c1 = 0), the square root f2m_sqrt
// (build_f2m.js:451-520, "Alg 9 adj" of eprint 2012/685) and b the twist's constant (bls12381/build_bls12381.js:49-52, bn128/build_bn128.js:45-48).
// One reference defect is not reproduced: g1m_LEMtoC tests infinity with the *Jacobian* predicate (g1m_isZero, :1175) on an
// affine input, i.e. it looks at the next point's x; here the affine predicate is used (as g1m_LEMtoU does, :1222).
#pragma once
#include "ec.cuh"

namespace b200 {

enum { CODEC_LEM_TO_U = 0, CODEC_LEM_TO_C = 1, CODEC_U_TO_LEM = 2, CODEC_C_TO_LEM = 3, CODEC_TO_AFFINE = 4, CODEC_TO_JACOBIAN = 5 };

// big-endian bytes <-> little-endian limbs (g1m__reverseBytes :1127-1164): word k of the output is the byte-swapped word N-1-k
template <class C> B200_DI void fe_store_be(uint8_t* p, const Fe<C::N>& a) {
  uint32_t* d = reinterpret_cast<uint32_t*>(p);
#pragma unroll
  for (int k = 0; k < C::N; k++) d[k] = __byte_perm(a.l[C::N - 1 - k], 0, 0x0123);
}
template <class C> B200_DI void fe_load_be(Fe<C::N>& a, const uint8_t* p) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
  for (int k = 0; k < C::N; k++) a.l[C::N - 1 - k] = __byte_perm(s[k], 0, 0x0123);
}
// f1m_sign (build_f1m.js:135-156) == -1  <=>  canonical value >= (q + 1) / 2
template <class C> B200_DI bool fe_is_greater_half_p(const Fe<C::N>& y_mont) {
  Fe<C::N> y; fe_from_mont_p<C>(y, y_mont);
  // (q + 1) / 2 limb by limb: q is odd, so (q + 1) / 2 = (q >> 1) + 1
  uint32_t h[C::N];
#pragma unroll
  for (int i = 0; i < C::N; i++) h[i] = (C::q(i) >> 1) | (i + 1 < C::N ? (C::q(i + 1) << 31) : 0u);
  uint32_t t, borrow;
  // y >= h + 1  <=>  y - h - 1 does not borrow
  sub_cc(t, y.l[0], h[0]);
#pragma unroll
  for (int i = 1; i < C::N; i++) subc_cc(t, y.l[i], h[i]);
  subc(borrow, 0, 0);
  if (borrow) return false;              // y < h
  // y >= h: greater-or-equal to h + 1 unless y == h
  uint32_t diff = 0;
#pragma unroll
  for (int i = 0; i < C::N; i++) diff |= y.l[i] ^ h[i];
  return diff != 0;
}
template <class C> B200_DI bool fe_is_greater_half(const Fe<C::N>& y_mont) {
  if constexpr (C::EXT == 2) {      // f2m_sign: decided by c1 unless it is zero
    using B = typename C::Base; Fe<B::N> y0, y1; fq2_get<C>(y0, y1, y_mont);
    return fe_is_zero<B>(y1) ? fe_is_greater_half_p<B>(y0) : fe_is_greater_half_p<B>(y1);
  } else return fe_is_greater_half_p<C>(y_mont);
}
// f1m_sqrt for q = 3 mod 4: a^((q+1)/4).  (build_f1m.js buildSqrt uses Tonelli-Shanks; for these fields it reduces to this power.
// The root returned may be either one -- callers pick by sign, as g1m_CtoLEM does.)
template <class C> __device__ __noinline__ void fe_sqrt_p(Fe<C::N>& r, const Fe<C::N>& a) {
  constexpr int N = C::N;
  uint32_t e[N + 1];
  // e = (q + 1) / 4 = (q >> 2) + 1   (q = 3 mod 4)
#pragma unroll
  for (int i = 0; i < N; i++) e[i] = (C::q(i) >> 2) | (i + 1 < N ? (C::q(i + 1) << 30) : 0u);
  e[0] += 1;                              // no carry: low two bits of q>>2 ... (checked for both fields in tests)
  Fe<N> acc; fe_set_one<C>(acc);
  for (int i = C::QBITS - 2; i >= 0; i--) {
    fe_sqr<C>(acc, acc);
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < N; k++) w = (k == (i >> 5)) ? e[k] : w;
    if ((w >> (i & 31)) & 1) fe_mul<C>(acc, acc, a);
  }
  r = acc;
}

// a^e in Fq2 for an exponent of the BASE field's size given as (q - SUB) >> SHIFT (f2m_exp with the constants of f2m_sqrt)
template <class C, int SUB, int SHIFT> __device__ __noinline__ void fq2_pow_q(Fe<C::N>& r, const Fe<C::N>& a) {
  using B = typename C::Base; constexpr int NB = B::N;
  uint32_t e[NB];
  { uint32_t t[NB]; uint32_t br = SUB;       // t = q - SUB (SUB is 1 or 3: only the low limb changes, q's low limb is larger)
#pragma unroll
    for (int i = 0; i < NB; i++) { t[i] = B::q(i) - (i == 0 ? br : 0u); }
#pragma unroll
    for (int i = 0; i < NB; i++) e[i] = (t[i] >> SHIFT) | (i + 1 < NB ? (t[i + 1] << (32 - SHIFT)) : 0u); }
  Fe<C::N> acc; fe_set_one<C>(acc);
  for (int i = B::QBITS - 1; i >= 0; i--) {
    fe_sqr<C>(acc, acc);
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < NB; k++) w = (k == (i >> 5)) ? e[k] : w;
    if ((w >> (i & 31)) & 1) fe_mul<C>(acc, acc, a);
  }
  r = acc;
}
// f2m_sqrt (build_f2m.js:451-520): a1 = a^((q-3)/4), alpha = a1^2 a, x0 = a1 a; alpha = -1: x = u x0, else x = (1 + alpha)^((q-1)/2) x0.
// (a non-square input -- conj(alpha) alpha = -1 -- traps in the reference; here it yields some field element, like the prime-field case)
template <class C> __device__ __noinline__ void fq2_sqrt(Fe<C::N>& r, const Fe<C::N>& a) {
  using B = typename C::Base;
  Fe<C::N> a1, alpha, x0, n1, one;
  fe_set_one<C>(one); fe_neg<C>(n1, one);
  fq2_pow_q<C, 3, 2>(a1, a);
  fe_sqr<C>(alpha, a1); fe_mul<C>(alpha, a, alpha);
  fe_mul<C>(x0, a1, a);
  if (fe_eq<C>(alpha, n1)) {          // multiply by u: (c0 + c1 u) u = -c1 + c0 u
    Fe<B::N> c0, c1; fq2_get<C>(c0, c1, x0); fe_neg_p<B>(c1, c1); fq2_put<C>(r, c1, c0);
  } else {
    Fe<C::N> b; fe_add<C>(b, one, alpha);
    fq2_pow_q<C, 1, 1>(b, b);
    fe_mul<C>(r, b, x0);
  }
}
template <class C> B200_DI void fe_sqrt(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) fq2_sqrt<C>(r, a); else fe_sqrt_p<C>(r, a);
}
// the curve constant b in Montgomery form: G1 y^2 = x^3 + 4 / + 3; G2 twists y^2 = x^3 + 4(1 + u) (BLS12-381) and x^3 + 3 / (9 + u) (BN254)
template <class C> B200_DI void codec_curve_b(Fe<C::N>& b) {
  if constexpr (C::EXT == 2) {
    using B = typename C::Base; Fe<B::N> b0, b1;
    if constexpr (B::N == 12) {
      Fe<B::N> one; fe_set_one<B>(one); fe_add_p<B>(b0, one, one); fe_add_p<B>(b0, b0, b0); b1 = b0;
    } else {
      const uint32_t k0[8] = {0x24a138e5u, 0x3267e6dcu, 0x59dbefa3u, 0xb5b4c5e5u, 0x1be06ac3u, 0x81be1899u, 0xceb8aaaeu, 0x2b149d40u};
      const uint32_t k1[8] = {0x85c315d2u, 0xe4a2bd06u, 0xe52d1852u, 0xa74fa084u, 0xeed8fdf4u, 0xcd2cafadu, 0x3af0fed4u, 0x009713b0u};
#pragma unroll
      for (int i = 0; i < 8; i++) { b0.l[i] = k0[i]; b1.l[i] = k1[i]; }
      fe_to_mont_p<B>(b0, b0); fe_to_mont_p<B>(b1, b1);
    }
    fq2_put<C>(b, b0, b1);
  } else {
    Fe<C::N> one; fe_set_one<C>(one); fe_add_p<C>(b, one, one);
    if constexpr (C::N == 12) fe_add_p<C>(b, b, b); else fe_add_p<C>(b, b, one);
  }
}

template <class C>
__global__ void __launch_bounds__(128) k_codec(int op, const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out) {
  constexpr int N = C::N, n8 = 4 * C::N;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op == CODEC_LEM_TO_U) {
    Affine<C> p; affine_load<C>(p, in, i);
    uint8_t* o = out + (uint64_t)i * 2 * n8;
    if (affine_is_inf<C>(p)) { fe_store_be<C>(o, p.x); fe_store_be<C>(o + n8, p.y); o[0] = 0x40; return; }
    fe_from_mont<C>(p.x, p.x); fe_from_mont<C>(p.y, p.y);
    fe_store_be<C>(o, p.x); fe_store_be<C>(o + n8, p.y);
  } else if (op == CODEC_LEM_TO_C) {
    Affine<C> p; affine_load<C>(p, in, i);
    uint8_t* o = out + (uint64_t)i * n8;
    if (affine_is_inf<C>(p)) { fe_store_be<C>(o, p.x); o[0] = 0x40; return; }
    bool greater = fe_is_greater_half<C>(p.y);
    fe_from_mont<C>(p.x, p.x);
    fe_store_be<C>(o, p.x);
    if (greater) o[0] |= 0x80;
  } else if (op == CODEC_U_TO_LEM) {
    const uint8_t* s = in + (uint64_t)i * 2 * n8;
    Affine<C> p;
    if (s[0] & 0x40) { affine_set_inf<C>(p); affine_store<C>(out, i, p); return; }
    fe_load_be<C>(p.x, s); fe_load_be<C>(p.y, s + n8);
    fe_to_mont<C>(p.x, p.x); fe_to_mont<C>(p.y, p.y);
    affine_store<C>(out, i, p);
  } else if (op == CODEC_C_TO_LEM) {
    const uint8_t* s = in + (uint64_t)i * n8;
    Affine<C> p;
    const uint32_t first = s[0];
    if (first & 0x40) { affine_set_inf<C>(p); affine_store<C>(out, i, p); return; }
    fe_load_be<C>(p.x, s);
    p.x.l[N - 1] &= 0x3fffffffu;                       // clear the two flag bits (top byte & 0x3F)
    fe_to_mont<C>(p.x, p.x);
    Fe<N> t, y, ny;
    fe_sqr<C>(t, p.x); fe_mul<C>(t, t, p.x);
    codec_curve_b<C>(y);                                           // b in Montgomery form
    fe_add<C>(t, t, y);
    fe_sqrt<C>(y, t);
    fe_neg<C>(ny, y);
    const bool y_greater = fe_is_greater_half<C>(y), want_greater = (first & 0x80) != 0;
    p.y = (y_greater == want_greater) ? y : ny;
    affine_store<C>(out, i, p);
  } else if (op == CODEC_TO_JACOBIAN) {
    Affine<C> p; affine_load<C>(p, in, i);
    char* o = reinterpret_cast<char*>(out) + (uint64_t)i * 3 * n8;
    Fe<N> one, z; fe_set_one<C>(one);
    if (affine_is_inf<C>(p)) { fe_set_zero<C>(z); fe_store<C>(o, z); fe_store<C>(o + n8, one); fe_store<C>(o + 2 * n8, z); return; }
    fe_store<C>(o, p.x); fe_store<C>(o + n8, p.y); fe_store<C>(o + 2 * n8, one);
  }
}

// g1m_batchToAffine: one inversion per thread amortised over GROUP points (Montgomery trick, build_batchinverse.js:4-140)
template <class C, int GROUP>
__global__ void __launch_bounds__(128) k_jacobian_to_affine(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out) {
  constexpr int n8 = 4 * C::N;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t base = (uint64_t)t * GROUP;
  if (base >= n) return;
  uint32_t cnt = (n - base < GROUP) ? (uint32_t)(n - base) : GROUP;
  Fe<C::N> pre[GROUP], acc; fe_set_one<C>(acc);
  for (uint32_t k = 0; k < cnt; k++) {
    Fe<C::N> z; fe_load<C>(z, in + (base + k) * 3 * n8 + 2 * n8);
    pre[k] = acc;
    if (!fe_is_zero<C>(z)) fe_mul<C>(acc, acc, z);
  }
  Fe<C::N> inv; fe_inv_fast<C>(inv, acc);
  for (int k = (int)cnt - 1; k >= 0; k--) {
    const uint8_t* s = in + (base + k) * 3 * n8;
    Fe<C::N> x, y, z; fe_load<C>(x, s); fe_load<C>(y, s + n8); fe_load<C>(z, s + 2 * n8);
    Affine<C> a;
    if (fe_is_zero<C>(z)) affine_set_inf<C>(a);
    else {
      Fe<C::N> zi, zi2, zi3; fe_mul<C>(zi, inv, pre[k]); fe_mul<C>(inv, inv, z);
      fe_sqr<C>(zi2, zi); fe_mul<C>(zi3, zi2, zi);
      fe_mul<C>(a.x, x, zi2); fe_mul<C>(a.y, y, zi3);
    }
    affine_store<C>(out, base + k, a);
  }
}

}  // namespace b200
