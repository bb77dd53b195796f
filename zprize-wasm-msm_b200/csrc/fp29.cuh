// fp29.cuh -- carry-free Montgomery multiplication in radix 2^29: a MEASURED ALTERNATIVE, not on the product path.
//
// Idea: hold elements as L limbs of 29 bits (L = 14 for BLS12-381, 9 for BN254), accumulate every partial product
// a_j*b_i (< 2^58) into a 64-bit column with a carry-less mad.wide, and resolve carries once per row: 2L^2 + L
// multiplies without any carry chain, instead of the 2N^2 + N carry-chained IMAD.WIDE.U32.X of fp.cuh.
// Montgomery radix is R' = 2^(29L); fe29_enter() / fe29_leave() change radix to and from the reference's R = 2^(32N).
//
// Result on B200 (tools/sweep.py --probe, bit-exact in tests/test_gpu_parity.py::test_radix29_multiplier_bit_exact):
//   radix 2^32 carry chains (fp.cuh) : 30.4e9 BLS12-381 multiplications/s = 99 % of the IMAD.WIDE.U32 peak
//   radix 2^29 carry-free (this file): 18.6e9 /s  (0.61x)
// Why it loses: IMAD.WIDE.U32 issues at 32 lanes/clk/SM whether or not it carries (9.18e12 /s; only the 32-bit IMAD
// reaches 64 lanes/clk/SM), so dropping the carry buys nothing, while ptxas splits every carry-less mad.wide into
// IMAD.WIDE (RZ addend) + IADD3 + IADD3.X, adding ~430 ALU instructions per multiplication.  The carry-chain
// multiplier is therefore the right form for this chip and stays on the hot path; this file documents the experiment.
#pragma once
#include "fp.cuh"

namespace b200 {

template <class C> struct R29 {
  static constexpr int L = (C::QBITS + 28) / 29;
  static constexpr uint32_t MASK = (1u << 29) - 1u;
  static constexpr uint32_t NP = C::NP & MASK;                           // -q^-1 mod 2^29
  __host__ __device__ static constexpr uint32_t q(int j) {              // limb j of q in radix 2^29
    const int bit = 29 * j, k = bit >> 5, r = bit & 31;
    const uint64_t lo = k < C::N ? C::q(k) : 0u, hi = (k + 1) < C::N ? C::q(k + 1) : 0u;
    return (uint32_t)(((lo | (hi << 32)) >> r) & MASK);
  }
};

template <class C> struct Fe29 { uint32_t l[R29<C>::L]; };

// packed N x 32-bit words (value < 2^(32N)) -> L limbs of 29 bits
template <class C> B200_DI void fe29_unpack(Fe29<C>& r, const Fe<C::N>& a) {
  constexpr int L = R29<C>::L, N = C::N;
#pragma unroll
  for (int j = 0; j < L; j++) {
    const int bit = 29 * j, k = bit >> 5, sh = bit & 31;
    uint32_t lo = a.l[k], hi = (k + 1 < N) ? a.l[k + 1] : 0u;
    r.l[j] = (sh + 29 <= 32 ? (lo >> sh) : __funnelshift_r(lo, hi, sh)) & R29<C>::MASK;
  }
}
// L normalised limbs (each < 2^29, value < 2^(32N)) -> packed words
template <class C> B200_DI void fe29_pack(Fe<C::N>& r, const Fe29<C>& a) {
  constexpr int L = R29<C>::L, N = C::N;
#pragma unroll
  for (int k = 0; k < N; k++) {
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < L; j++) {
      const int d = 29 * j - 32 * k;                 // position of limb j relative to word k
      if (d > -29 && d < 32) w |= (d >= 0) ? (a.l[j] << d) : (a.l[j] >> (-d));
    }
    r.l[k] = w;
  }
}

// r = a - q if a >= q else a ; limbs normalised in and out; input value < 2q
template <class C> B200_DI void fe29_reduce_once(Fe29<C>& a) {
  constexpr int L = R29<C>::L;
  uint32_t d[L]; int32_t c = 0;
#pragma unroll
  for (int j = 0; j < L; j++) {
    int32_t v = (int32_t)a.l[j] - (int32_t)R29<C>::q(j) + c;
    d[j] = (uint32_t)v & R29<C>::MASK; c = v >> 29;
  }
#pragma unroll
  for (int j = 0; j < L; j++) a.l[j] = c ? a.l[j] : d[j];      // c = -1: borrow out, a < q
}

B200_DI void mad_wide(uint64_t& acc, uint32_t a, uint32_t b) { asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b)); }

// r = a * b / R' mod q.  Inputs: limbs < 2^29 + 2^8 (normalised or weakly normalised), values < 8q.
// Output: limbs normalised, value < q (CANON) or < q + 2^-7 q (not CANON: one conditional subtraction skipped).
template <class C, bool CANON = true>
B200_DI void fe29_mul(Fe29<C>& r, const Fe29<C>& a, const Fe29<C>& b) {
  constexpr int L = R29<C>::L;
  uint64_t t[L + 1];
#pragma unroll
  for (int j = 0; j <= L; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
#pragma unroll
    for (int j = 0; j < L; j++) mad_wide(t[j], a.l[j], b.l[i]);
    const uint32_t m = ((uint32_t)t[0] * R29<C>::NP) & R29<C>::MASK;
#pragma unroll
    for (int j = 0; j < L; j++) mad_wide(t[j], m, R29<C>::q(j));
    t[1] += t[0] >> 29;                              // low 29 bits of t[0] are zero now: retire the column
#pragma unroll
    for (int j = 0; j < L; j++) t[j] = t[j + 1];
    t[L] = 0;
  }
#pragma unroll
  for (int j = 0; j < L - 1; j++) { r.l[j] = (uint32_t)t[j] & R29<C>::MASK; t[j + 1] += t[j] >> 29; }
  r.l[L - 1] = (uint32_t)t[L - 1];
  if (CANON) fe29_reduce_once<C>(r);
}

// change of Montgomery radix at the engine boundary
template <class C> struct R29K {
  // (R'^2 / R) mod q as packed words: enter(X) = X * this / R' = x R'
  __host__ __device__ static constexpr uint32_t enter(int i);
};
template <> __host__ __device__ constexpr uint32_t R29K<BLS12_381>::enter(int i) {
  constexpr uint32_t t[12] = {0x9fddebbdu, 0x6749ea8eu, 0x9e0a47ceu, 0xd2ca681du, 0x5794f6cau, 0xa09f3699u,
                              0xe3563d64u, 0xc52a8410u, 0xdfabd89eu, 0xb258369au, 0x17e957b5u, 0x17326359u};
  return t[i];
}
template <> __host__ __device__ constexpr uint32_t R29K<BN254>::enter(int i) {
  constexpr uint32_t t[8] = {0x13349ca1u, 0xb34bb095u, 0xf028f972u, 0x1e880124u, 0xa6092b95u, 0xe56cdd25u, 0xdce9ed32u, 0x05800320u};
  return t[i];
}
// x*R (packed, canonical) -> x*R' (limbs, canonical)
template <class C> B200_DI void fe29_enter(Fe29<C>& r, const Fe<C::N>& a) {
  Fe<C::N> kw; Fe29<C> k, u;
#pragma unroll
  for (int i = 0; i < C::N; i++) kw.l[i] = R29K<C>::enter(i);
  fe29_unpack<C>(k, kw); fe29_unpack<C>(u, a);
  fe29_mul<C, true>(r, u, k);
}
// x*R' (limbs) -> x*R (packed, canonical): multiply by (R mod q) / R'
template <class C> B200_DI void fe29_leave(Fe<C::N>& r, const Fe29<C>& a) {
  Fe<C::N> kw; Fe29<C> k, u;
#pragma unroll
  for (int i = 0; i < C::N; i++) kw.l[i] = C::one(i);
  fe29_unpack<C>(k, kw);
  fe29_mul<C, true>(u, a, k);
  fe29_pack<C>(r, u);
}

// throughput probe: ITER dependent radix-2^29 multiplications per thread
template <class C>
__global__ void __launch_bounds__(256) k_fpmul29_probe(uint32_t iters, const void* __restrict__ in, void* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe<C::N> xw; fe_load<C>(xw, reinterpret_cast<const char*>(in) + (uint64_t)(i & 1023) * 4 * C::N);
  Fe29<C> x, y; fe29_unpack<C>(x, xw);
#pragma unroll
  for (int j = 0; j < R29<C>::L; j++) x.l[j] &= R29<C>::MASK;
  x.l[R29<C>::L - 1] &= 0xff;
  y = x;
  for (uint32_t k = 0; k < iters; k++) fe29_mul<C, false>(y, y, x);
  fe29_pack<C>(xw, y);
  fe_store<C>(reinterpret_cast<char*>(out) + (uint64_t)i * 4 * C::N, xw);
}
// parity hook: packed canonical a, b -> packed canonical a*b/R' (pure radix-2^29 Montgomery product)
template <class C>
__global__ void k_fp29_mul(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ r, uint32_t n, int mode) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<C::N> x, y, z; Fe29<C> u, v, w;
  fe_load<C>(x, reinterpret_cast<const char*>(a) + (uint64_t)i * 4 * C::N);
  fe_load<C>(y, reinterpret_cast<const char*>(b) + (uint64_t)i * 4 * C::N);
  if (mode == 0) { fe29_unpack<C>(u, x); fe29_unpack<C>(v, y); fe29_mul<C, true>(w, u, v); fe29_pack<C>(z, w); }
  else { fe29_enter<C>(u, x); fe29_enter<C>(v, y); fe29_mul<C, true>(w, u, v); fe29_leave<C>(z, w); }   // == f1m_mul in the reference's radix
  fe_store<C>(reinterpret_cast<char*>(r) + (uint64_t)i * 4 * C::N, z);
}

}  // namespace b200
