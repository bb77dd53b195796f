// glv.cuh -- GLV pre-pass for BLS12-381 G1 on the GPU.
//
// GPU form of wasmcurves/src/build_glv.js (BLS12-381 only, like the reference):
//   g1m_glv_decomposeScalar          :53-146   k -> (|k1|, |k2|, sign), k = k1 + k2*lambda-style split with the reference's lattice basis
//   g1m_glv_endomorphism             :150-174  (x, y) -> (beta*x, +-y)
//   g1m_glv_preprocessEndomorphism   :178-263  N points / N scalars -> 2N points / 2N scalars (each < 2^128, stored in 32 bytes)
// The reference works in 512-bit integers with a generic long division (int512_div); the divisor is the constant r, so
// here the two quotients come from two comparisons (q1 = floor(k / r) is 0, 1 or 2 for a 256-bit k) and one Barrett
// estimate with a single correction (q2 = floor(k * (-v1) / r)).  One thread per scalar, plain 32-bit limb loops: the
// pre-pass is a few hundred integer instructions per point, noise next to the MSM itself.
// Results are bit-identical to the reference's: the low 128 bits of |k1| and |k2| and the two sign bits
// (bit 0: k1 >= 0, bit 1: k2 >= 0; build_glv.js:117-137), tested against test/glv.js:50-65,103-192.
#pragma once
#include "ec.cuh"

namespace b200 {

__device__ __constant__ uint32_t GLV_R[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};      // build_glv.js:22 (divisor = r)
__device__ __constant__ uint32_t GLV_2R[8] = {0x00000002u, 0xfffffffeu, 0xfffcb7fdu, 0xa77b4805u, 0x1343b00au, 0x6673b010u, 0x533afa90u, 0xe7db4ea6u};
__device__ __constant__ uint32_t GLV_NEGV1[4] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u};                                                      // build_glv.js:19
__device__ __constant__ uint32_t GLV_U0[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};                                                         // build_glv.js:17
__device__ __constant__ uint32_t GLV_MU[9] = {0x0c0d6393u, 0x42737a02u, 0xbe4bad71u, 0x65043eb4u, 0x07e08ed3u, 0x38b5dcb7u, 0xfede377cu, 0x355094edu, 0x00000002u};   // floor(2^512 / r)
// beta * 2^384 mod q (build_glv.js:21,29): a primitive cube root of unity in Fq, Montgomery form
__device__ __constant__ uint32_t GLV_BETA_M[12] = {0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au, 0x74fd029bu, 0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu};

// r[0..NA+NB) = a[0..NA) * b[0..NB)
template <int NA, int NB> B200_DI void limbs_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 0; i < NA + NB; i++) r[i] = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) { c += (uint64_t)a[i] * b[j] + r[i + j]; r[i + j] = (uint32_t)c; c >>= 32; }
    r[i + NB] = (uint32_t)c;
  }
}
// a[0..N) -= b[0..NB) (b zero-extended), two's complement wrap-around
template <int N, int NB> B200_DI void limbs_sub(uint32_t* a, const uint32_t* b) {
  int64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { c += (int64_t)a[i] - (int64_t)(i < NB ? b[i] : 0u); a[i] = (uint32_t)c; c >>= 32; }
}
template <int N> B200_DI bool limbs_gte(const uint32_t* a, const uint32_t* b) {      // a >= b
#pragma unroll
  for (int i = N - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return true;
}
template <int N> B200_DI void limbs_negate(uint32_t* a) {
  uint64_t c = 1;
#pragma unroll
  for (int i = 0; i < N; i++) { c += (uint32_t)~a[i]; a[i] = (uint32_t)c; c >>= 32; }
}

// k (8 limbs) -> k1abs, k2abs (4 limbs each), returns sign (bit 0: k1 >= 0, bit 1: k2 >= 0)
B200_DI uint32_t glv_decompose(const uint32_t (&k)[8], uint32_t (&k1abs)[4], uint32_t (&k2abs)[4]) {
  uint32_t rr[8], r2[8], nv[4], u0[4], mu[9];
#pragma unroll
  for (int i = 0; i < 8; i++) { rr[i] = GLV_R[i]; r2[i] = GLV_2R[i]; }
#pragma unroll
  for (int i = 0; i < 4; i++) { nv[i] = GLV_NEGV1[i]; u0[i] = GLV_U0[i]; }
#pragma unroll
  for (int i = 0; i < 9; i++) mu[i] = GLV_MU[i];
  // q1 = floor(k / r)                                                  (build_glv.js:104)
  const uint32_t q1 = (limbs_gte<8>(k, rr) ? 1u : 0u) + (limbs_gte<8>(k, r2) ? 1u : 0u);
  // q2 = floor(k * (-v1) / r): Barrett estimate from floor(2^512 / r), off by at most one      (build_glv.js:106-107)
  uint32_t x[12], t[21], q2[5], p[13], rem[9];
  limbs_mul<8, 4>(x, k, nv);
  limbs_mul<12, 9>(t, x, mu);
#pragma unroll
  for (int i = 0; i < 5; i++) q2[i] = t[16 + i];
  limbs_mul<5, 8>(p, q2, rr);
#pragma unroll
  for (int i = 0; i < 9; i++) rem[i] = x[i];
  limbs_sub<9, 9>(rem, p);                                              // exact remainder < 2r fits in 9 limbs
  { uint32_t r9[9];
#pragma unroll
    for (int i = 0; i < 8; i++) r9[i] = rr[i];
    r9[8] = 0;
    if (limbs_gte<9>(rem, r9)) { uint64_t c = 1;
#pragma unroll
      for (int i = 0; i < 5; i++) { c += q2[i]; q2[i] = (uint32_t)c; c >>= 32; } } }
  // k1 = k - q1*v0 - q2*u0, v0 = 1                                     (build_glv.js:110-112)
  uint32_t a[10], m[9], one[1] = {q1};
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = k[i];
  a[8] = a[9] = 0;
  limbs_sub<10, 1>(a, one);
  limbs_mul<5, 4>(m, q2, u0);
  limbs_sub<10, 9>(a, m);
  // k2 = -q1*v1 - q2*u1, u1 = 1, -v1 = nv                              (build_glv.js:115-117)
  uint32_t b[6], qq[1] = {q1};
  { uint32_t bb[5]; limbs_mul<1, 4>(bb, qq, nv);
#pragma unroll
    for (int i = 0; i < 5; i++) b[i] = bb[i];
    b[5] = 0; }
  limbs_sub<6, 5>(b, q2);
  uint32_t sign = 0;
  if (!(a[9] >> 31)) sign |= 1u; else limbs_negate<10>(a);             // isPositive: top bit clear, zero counts as positive (:33-50)
  if (!(b[5] >> 31)) sign |= 2u; else limbs_negate<6>(b);
#pragma unroll
  for (int i = 0; i < 4; i++) { k1abs[i] = a[i]; k2abs[i] = b[i]; }     // the reference keeps the low two 64-bit words (:133-136)
  return sign;
}

// scalars: n x 32 bytes -> out_scalars: n x 64 bytes (|k1| in bytes 0..15, |k2| in bytes 32..47, rest zero), out_signs: n words (nullable)
__global__ void __launch_bounds__(128) k_glv_decompose(const uint32_t* __restrict__ scalars, uint32_t n, uint32_t* __restrict__ out_scalars, uint32_t* __restrict__ out_signs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8], k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 8; j++) k[j] = scalars[(uint64_t)i * 8 + j];
  const uint32_t sign = glv_decompose(k, k1, k2);
  uint4* o = reinterpret_cast<uint4*>(out_scalars + (uint64_t)i * 16);
  o[0] = make_uint4(k1[0], k1[1], k1[2], k1[3]); o[1] = make_uint4(0, 0, 0, 0);
  o[2] = make_uint4(k2[0], k2[1], k2[2], k2[3]); o[3] = make_uint4(0, 0, 0, 0);
  if (out_signs) out_signs[i] = sign;
}

// points: n affine Montgomery points; signs from k_glv_decompose -> out_points: 2n points
//   out[2i] = (x, s0 ? y : -y), out[2i+1] = (beta*x, s1 ? y : -y)        (build_glv.js:150-174, 213-257)
__global__ void __launch_bounds__(128) k_glv_points(const void* __restrict__ points, const uint32_t* __restrict__ signs, uint32_t n, void* __restrict__ out_points) {
  using C = BLS12_381;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> p; affine_load<C>(p, points, i);
  const uint32_t sign = signs[i];
  Fe<C::N> beta, ny;
#pragma unroll
  for (int j = 0; j < C::N; j++) beta.l[j] = GLV_BETA_M[j];
  fe_neg<C>(ny, p.y);
  Affine<C> a, b;
  a.x = p.x; a.y = (sign & 1u) ? p.y : ny;
  fe_mul<C>(b.x, p.x, beta); b.y = (sign & 2u) ? p.y : ny;
  affine_store<C>(out_points, 2ull * i, a);
  affine_store<C>(out_points, 2ull * i + 1, b);
}

}  // namespace b200
