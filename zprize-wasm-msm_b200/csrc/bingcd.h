// bingcd.h -- modular inversion by the optimised binary GCD with 31-step inner loops on 64-bit approximations
// (T. Pornin, "Optimized Binary GCD for Modular Inversion", 2020, algorithm 2, k = 32).
//
// f1m_inverse in the reference is an extended Euclid on full-length integers (wasmcurves/src/build_int.js:922-1064 behind
// build_f1m.js:1112-1122).  On the GPU the inversion sits at the root of every batch inversion (f1m_batchInverse,
// build_batchinverse.js:90): ONE thread runs it while a whole tree round waits, so its latency -- not its throughput -- matters.
// The plain binary Euclid (fe_inv_fast_euclid_p in fp.cuh) does ~2*log2(q) full-width shift/subtract steps; here 31 steps at a time
// run on two 64-bit words (the low 31 bits and the top 33 bits of a and b), the four 32-bit update factors are then applied to the
// full-length values once: ceil((2*QBITS-1)/31) outer rounds (25 for BLS12-381 Fq, 17 for the 254/255-bit fields), ~5x fewer instructions.
//
// Plain C++ (no PTX): the same code is compiled by g++ for tests/host_inv_harness.cpp and by nvcc for the device.
// All loops have compile-time bounds and constant indices so that the limb arrays stay in registers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HDI __host__ __device__ __forceinline__
#else
#define B200_HDI inline
#endif

namespace b200 {

B200_HDI int bingcd_clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
B200_HDI uint32_t bingcd_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

// out (N+1 limbs) = x * f, f <= 2^31
template <int N> B200_HDI void bingcd_mul1(uint32_t (&out)[N + 1], const uint32_t (&x)[N], uint32_t f) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { c += (uint64_t)x[i] * f; out[i] = (uint32_t)c; c >>= 32; }
  out[N] = (uint32_t)c;
}
// r = |x*f + y*g| >> 31 for signed factors |f|, |g| <= 2^31 (the sum is an exact multiple of 2^31); returns true when x*f + y*g < 0
template <int N> B200_HDI bool bingcd_lincomb(uint32_t (&r)[N], const uint32_t (&x)[N], int64_t f, const uint32_t (&y)[N], int64_t g) {
  const bool nf = f < 0, ng = g < 0;
  uint32_t P[N + 1], Q[N + 1];
  bingcd_mul1<N>(P, x, (uint32_t)(nf ? -f : f));
  bingcd_mul1<N>(Q, y, (uint32_t)(ng ? -g : g));
  uint32_t D[N + 1]; bool neg;
  if (nf == ng) {                      // same sign: magnitudes add
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i <= N; i++) { c += (uint64_t)P[i] + Q[i]; D[i] = (uint32_t)c; c >>= 32; }
    neg = nf;
  } else {                             // opposite signs: D = P - Q, negated when it borrows
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i <= N; i++) { c += (int64_t)P[i] - (int64_t)Q[i]; D[i] = (uint32_t)c; c >>= 32; }
    const bool borrow = c < 0;
    if (borrow) {
      uint64_t k = 1;
#pragma unroll
      for (int i = 0; i <= N; i++) { k += (uint32_t)~D[i]; D[i] = (uint32_t)k; k >>= 32; }
    }
    neg = nf != borrow;                // P - Q >= 0 has the sign of f
  }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = (D[i] >> 31) | (D[i + 1] << 1);
  return neg;
}
// r = (u*f + v*g) / 2^31 mod m for u, v in [0, m); signed factors |f|, |g| <= 2^31
template <class C> B200_HDI void bingcd_lincomb_mod(uint32_t (&r)[C::N], const uint32_t (&u)[C::N], int64_t f, const uint32_t (&v)[C::N], int64_t g) {
  constexpr int N = C::N;
  uint32_t uf[N], vg[N];
  {   // a negative factor is moved onto the residue: f*u = |f| * (m - u) mod m
    int64_t bu = 0, bv = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      bu += (int64_t)C::q(i) - (int64_t)u[i]; const uint32_t nu = (uint32_t)bu; bu >>= 32;
      bv += (int64_t)C::q(i) - (int64_t)v[i]; const uint32_t nv = (uint32_t)bv; bv >>= 32;
      uf[i] = f < 0 ? nu : u[i]; vg[i] = g < 0 ? nv : v[i];
    }
  }
  const uint32_t fa = (uint32_t)(f < 0 ? -f : f), ga = (uint32_t)(g < 0 ? -g : g);
  uint32_t t[N + 1];
  { uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) { const uint64_t p = (uint64_t)uf[i] * fa, q = (uint64_t)vg[i] * ga;
      c += (uint32_t)p; c += (uint32_t)q; t[i] = (uint32_t)c; c = (c >> 32) + (p >> 32) + (q >> 32); }
    t[N] = (uint32_t)c; }
  // t + k*m = 0 mod 2^31  (k < 2^31), then shift: the result is < 3m
  const uint32_t k = (t[0] * C::NP) & 0x7fffffffu;
  { uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) { c += (uint64_t)C::q(i) * k + t[i]; t[i] = (uint32_t)c; c >>= 32; }
    t[N] += (uint32_t)c; }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = (t[i] >> 31) | (t[i + 1] << 1);
  const uint32_t top = t[N] >> 31;     // bit 32N of the shifted value (only while it is >= m)
  uint32_t hi = top;
#pragma unroll
  for (int rep = 0; rep < 2; rep++) {
    uint32_t d[N]; int64_t b = 0;
#pragma unroll
    for (int i = 0; i < N; i++) { b += (int64_t)r[i] - (int64_t)C::q(i); d[i] = (uint32_t)b; b >>= 32; }
    const bool ge = hi != 0 || b == 0;   // r >= m
    if (ge) {
      if (b != 0) hi -= 1;               // the borrow consumes the extra top bit
#pragma unroll
      for (int i = 0; i < N; i++) r[i] = d[i];
    }
  }
}

// r = y^-1 mod q as a plain residue (y any residue in [0, q); 0 -> 0).  q = the modulus of field class C.
template <class C> B200_HDI void bingcd_inverse(uint32_t (&r)[C::N], const uint32_t (&y)[C::N]) {
  constexpr int N = C::N;
  constexpr int ROUNDS = (2 * C::QBITS - 1 + 30) / 31;
  uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = y[i]; b[i] = C::q(i); u[i] = (i == 0); v[i] = 0; }
#pragma unroll 1
  for (int round = 0; round < ROUNDS; round++) {
    // ---- 64-bit approximations: low 31 bits + the 33 bits below the common top (exact when both fit in 64 bits)
    int t = 1; uint32_t topw = 0;
#pragma unroll
    for (int i = N - 1; i >= 2; i--) { const uint32_t c = a[i] | b[i]; if (topw == 0 && c != 0) { topw = c; t = i; } }
    uint32_t orl = 0;
#pragma unroll
    for (int i = 0; i < N; i++) orl |= a[i];
    if (orl == 0) break;                                   // a == 0: b = gcd, v is final
    uint64_t xa, xb;
    if (topw == 0) { xa = ((uint64_t)a[1] << 32) | a[0]; xb = ((uint64_t)b[1] << 32) | b[0]; }
    else {
      uint32_t at = 0, at1 = 0, at2 = 0, bt = 0, bt1 = 0, bt2 = 0;
#pragma unroll
      for (int i = 2; i < N; i++) if (i == t) { at = a[i]; at1 = a[i - 1]; at2 = a[i - 2]; bt = b[i]; bt1 = b[i - 1]; bt2 = b[i - 2]; }
      const int s = 32 - bingcd_clz32(topw);               // bits used in the top limb: 1..32
      const uint64_t wa = ((uint64_t)at << 32) | at1, wb = ((uint64_t)bt << 32) | bt1;
      const uint64_t ha = s == 32 ? wa : ((wa << (32 - s)) | ((uint64_t)at2 >> s));
      const uint64_t hb = s == 32 ? wb : ((wb << (32 - s)) | ((uint64_t)bt2 >> s));
      xa = (ha & 0xffffffff80000000ull) | (a[0] & 0x7fffffffu);
      xb = (hb & 0xffffffff80000000ull) | (b[0] & 0x7fffffffu);
    }
    // ---- 31 binary-GCD steps on the approximations, recording the update matrix
    int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
    for (int j = 0; j < 31; j++) {
      const bool odd = (xa & 1) != 0;
      const bool sw = odd && xa < xb;
      if (sw) { const uint64_t tx = xa; xa = xb; xb = tx; int64_t tf = f0; f0 = f1; f1 = tf; tf = g0; g0 = g1; g1 = tf; }
      if (odd) { xa -= xb; f0 -= f1; g0 -= g1; }
      xa >>= 1; f1 <<= 1; g1 <<= 1;
    }
    // ---- apply it to the full-length values
    uint32_t na[N], nb[N];
    const bool nega = bingcd_lincomb<N>(na, a, f0, b, g0);
    const bool negb = bingcd_lincomb<N>(nb, a, f1, b, g1);
    if (nega) { f0 = -f0; g0 = -g0; }
    if (negb) { f1 = -f1; g1 = -g1; }
    uint32_t nu[N], nv[N];
    bingcd_lincomb_mod<C>(nu, u, f0, v, g0);
    bingcd_lincomb_mod<C>(nv, u, f1, v, g1);
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
  }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = v[i];
}

}  // namespace b200
