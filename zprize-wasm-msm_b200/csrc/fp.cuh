// fp.cuh -- Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces (on the GPU) the reference's generated WASM field layer:
//   wasmcurves/src/build_f1m.js:71-105 (add/sub), :466-777 (CIOS mul), :779-1076 (square),
//   wasmcurves/src/build_int.js:148-279 (gte/add/sub on limbs).
// Same value semantics: elements are n32 little-endian u32 limbs in Montgomery form a*R mod q,
// R = 2^(32*n32), always fully reduced (< q) at every function boundary in this file.
//
// The multiplier is written for the Blackwell integer pipe: every 32x32->64 limb product is a
// mad.lo.cc/madc.hi.cc pair that ptxas fuses into ONE IMAD.WIDE.U32 with carry in/out, and the
// partial products are kept in two interleaved accumulators (even / odd limb alignment) so that a
// whole row a*b_i is two independent carry chains of 64-bit adds -- no per-product carry fix-up.
// Count per multiplication: 2*N*N + N limb products (300 for BLS12-381, 136 for BN254), the same
// algorithmic figure the reference's CIOS has (build_f1m.js:575-660).
#pragma once
#include <stdint.h>
#include "field_params.h"
#include "bingcd.h"

namespace b200 {

// Quadratic extension Fq2 = Fq[u]/(u^2 + 1): the coordinate field of G2 on both curves (build_f2m.js; build_bls12381.js:48 and
// build_bn128.js:44 pass f1m_neg as the multiplication by the non-residue).  An element c0 + c1*u is stored c0 || c1 (the f2m layout)
// and is, to every template below and to every kernel, ONE field element of N = 2*Nb limbs: the fe_* entry points dispatch on EXT,
// so the whole G1 pipeline (sort, batch-affine tree, fold, window tables, batches) instantiates unchanged for G2.
template <class B> struct Fq2 {
  using Base = B;
  static constexpr int ID = B::ID + 2;
  static constexpr int EXT = 2;
  static constexpr int N = 2 * B::N;
  static constexpr int QBITS = B::QBITS;
  __host__ __device__ static constexpr uint32_t one(int i) { return i < B::N ? B::one(i) : 0u; }      // (R mod q) + 0*u
};

template <int N> struct alignas(16) Fe { uint32_t l[N]; };
// non-template kernels defined in headers: internal linkage in the translation units that only borrow the headers for their templates (tree.cu)
#if defined(TREE_CURVE)
#define B200_KERNEL static __global__
#else
#define B200_KERNEL __global__
#endif
#ifndef B200_DI
#define B200_DI __device__ __forceinline__
#endif
// field interface (dispatching on C::EXT; the *_p functions below are the prime-field implementations)
template <class C> B200_DI void fe_add(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b);
template <class C> B200_DI void fe_sub(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b);
template <class C> B200_DI void fe_neg(Fe<C::N>& r, const Fe<C::N>& a);
template <class C> B200_DI void fe_mul(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b);
template <class C> B200_DI void fe_sqr(Fe<C::N>& r, const Fe<C::N>& a);
template <class C> B200_DI void fe_to_mont(Fe<C::N>& r, const Fe<C::N>& a);
template <class C> B200_DI void fe_from_mont(Fe<C::N>& r, const Fe<C::N>& a);
template <class C> B200_DI void fe_inv(Fe<C::N>& r, const Fe<C::N>& a);
template <class C> B200_DI void fe_inv_fast(Fe<C::N>& r, const Fe<C::N>& a);

// ------------------------------------------------------------------ PTX carry-chain primitives
// asm volatile keeps the statements in program order; the condition-code register is only written
// by these instructions, so a chain spread over several statements is safe.
B200_DI uint32_t ptx_mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DI uint32_t ptx_mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DI void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DI void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DI void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DI void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DI void add_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DI void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DI void addc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DI void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DI void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DI void subc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }

// ------------------------------------------------------------------ basic predicates / moves
template <class C> B200_DI bool fe_is_zero(const Fe<C::N>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < C::N; i++) o |= a.l[i];
  return o == 0;
}
template <class C> B200_DI bool fe_eq(const Fe<C::N>& a, const Fe<C::N>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < C::N; i++) o |= a.l[i] ^ b.l[i];
  return o == 0;
}
template <class C> B200_DI void fe_set_zero(Fe<C::N>& a) {
#pragma unroll
  for (int i = 0; i < C::N; i++) a.l[i] = 0;
}
template <class C> B200_DI void fe_set_one(Fe<C::N>& a) {
#pragma unroll
  for (int i = 0; i < C::N; i++) a.l[i] = C::one(i);
}

// r = a - q if a >= q else a   (input < 2q)
template <class C> B200_DI void fe_reduce_once(Fe<C::N>& a) {
  constexpr int N = C::N;
  uint32_t t[N], borrow;
  sub_cc(t[0], a.l[0], C::q(0));
#pragma unroll
  for (int i = 1; i < N; i++) subc_cc(t[i], a.l[i], C::q(i));
  subc(borrow, 0, 0);                 // 0 if a >= q, 0xffffffff otherwise
#pragma unroll
  for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : t[i];
}

// f1m_add (build_f1m.js:71-89)
template <class C> B200_DI void fe_add_p(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  constexpr int N = C::N;
  add_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], a.l[i], b.l[i]);
  addc(r.l[N - 1], a.l[N - 1], b.l[N - 1]);      // q < 2^(32N-1): a + b < 2q never carries out
  fe_reduce_once<C>(r);
}

// f1m_sub (build_f1m.js:91-105)
template <class C> B200_DI void fe_sub_p(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  constexpr int N = C::N;
  uint32_t borrow;
  sub_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N; i++) subc_cc(r.l[i], a.l[i], b.l[i]);
  subc(borrow, 0, 0);
  // add q back under mask
  add_cc(r.l[0], r.l[0], C::q(0) & borrow);
#pragma unroll
  for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], r.l[i], C::q(i) & borrow);
  addc(r.l[N - 1], r.l[N - 1], C::q(N - 1) & borrow);
}

template <class C> B200_DI void fe_neg_p(Fe<C::N>& r, const Fe<C::N>& a) {
  constexpr int N = C::N;
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < N; i++) nz |= a.l[i];
  uint32_t t[N];
  sub_cc(t[0], C::q(0), a.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) subc_cc(t[i], C::q(i), a.l[i]);
  subc(t[N - 1], C::q(N - 1), a.l[N - 1]);
#pragma unroll
  for (int i = 0; i < N; i++) r.l[i] = nz ? t[i] : 0u;
}

template <class C> B200_DI void fe_dbl(Fe<C::N>& r, const Fe<C::N>& a) { fe_add<C>(r, a, a); }

// ------------------------------------------------------------------ Montgomery multiplication
// One CIOS step on the two interleaved accumulators.  Frame: value = sum E[k] 2^(32k) + sum O[k] 2^(32(k+1)).
// On entry (not first) O still holds the previous step's even accumulator (low limb zero), which is
// consumed with a two-limb right shift while the odd-limb products of this row are added.
template <class C, bool FIRST>
B200_DI void mont_row(uint32_t (&E)[C::N], uint32_t (&O)[C::N], const uint32_t (&a)[C::N], uint32_t bi) {
  constexpr int N = C::N;
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) { E[j] = ptx_mul_lo(a[j], bi); E[j + 1] = ptx_mul_hi(a[j], bi); }
#pragma unroll
    for (int j = 1; j < N; j += 2) { O[j - 1] = ptx_mul_lo(a[j], bi); O[j] = ptx_mul_hi(a[j], bi); }
  } else {
    add_cc(E[0], E[0], O[1]);
#pragma unroll
    for (int j = 1; j < N - 1; j += 2) { madc_lo_cc(O[j - 1], a[j], bi, O[j + 1]); madc_hi_cc(O[j], a[j], bi, O[j + 2]); }
    madc_lo_cc(O[N - 2], a[N - 1], bi, 0);
    madc_hi(O[N - 1], a[N - 1], bi, 0);
    mad_lo_cc(E[0], a[0], bi, E[0]);
    madc_hi_cc(E[1], a[0], bi, E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) { madc_lo_cc(E[j], a[j], bi, E[j]); madc_hi_cc(E[j + 1], a[j], bi, E[j + 1]); }
    addc(O[N - 1], O[N - 1], 0);
  }
  uint32_t m = E[0] * C::NP;
  mad_lo_cc(O[0], C::q(1), m, O[0]);
  madc_hi_cc(O[1], C::q(1), m, O[1]);
#pragma unroll
  for (int j = 3; j < N; j += 2) { madc_lo_cc(O[j - 1], C::q(j), m, O[j - 1]); madc_hi_cc(O[j], C::q(j), m, O[j]); }
  mad_lo_cc(E[0], C::q(0), m, E[0]);
  madc_hi_cc(E[1], C::q(0), m, E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) { madc_lo_cc(E[j], C::q(j), m, E[j]); madc_hi_cc(E[j + 1], C::q(j), m, E[j + 1]); }
  addc(O[N - 1], O[N - 1], 0);
}

// f1m_mul (build_f1m.js:466-777): r = a*b/R mod q, fully reduced.  r may alias a or b.  Interleaved (CIOS) form.
template <class C> B200_DI void fe_mul_cios(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  constexpr int N = C::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t E[N], O[N];
  mont_row<C, true>(E, O, a.l, b.l[0]);
  mont_row<C, false>(O, E, a.l, b.l[1]);
#pragma unroll
  for (int i = 2; i < N; i += 2) {
    mont_row<C, false>(E, O, a.l, b.l[i]);
    mont_row<C, false>(O, E, a.l, b.l[i + 1]);
  }
  // after an even number of rows: E = live even accumulator, O = stale even (low limb zero), pending 1-limb shift
  Fe<N> t;
  add_cc(t.l[0], E[0], O[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(t.l[k], E[k], O[k + 1]);
  addc(t.l[N - 1], E[N - 1], 0);
  fe_reduce_once<C>(t);
  r = t;
}

// Two independent multiplications with their rows interleaved in program order: r1 = a1*b1, r2 = a2*b2.  The carry chains of
// one multiplication are serial (asm volatile keeps them in order), so a single multiplication exposes only the parallelism of
// its two accumulators; interleaving a second, independent one doubles the work available between dependent instructions.
template <class C> B200_DI void fe_mul2(Fe<C::N>& r1, const Fe<C::N>& a1, const Fe<C::N>& b1, Fe<C::N>& r2, const Fe<C::N>& a2, const Fe<C::N>& b2) {
  constexpr int N = C::N;
  uint32_t E1[N], O1[N], E2[N], O2[N];
  mont_row<C, true>(E1, O1, a1.l, b1.l[0]);
  mont_row<C, true>(E2, O2, a2.l, b2.l[0]);
  mont_row<C, false>(O1, E1, a1.l, b1.l[1]);
  mont_row<C, false>(O2, E2, a2.l, b2.l[1]);
#pragma unroll
  for (int i = 2; i < N; i += 2) {
    mont_row<C, false>(E1, O1, a1.l, b1.l[i]);
    mont_row<C, false>(E2, O2, a2.l, b2.l[i]);
    mont_row<C, false>(O1, E1, a1.l, b1.l[i + 1]);
    mont_row<C, false>(O2, E2, a2.l, b2.l[i + 1]);
  }
  Fe<N> t1, t2;
  add_cc(t1.l[0], E1[0], O1[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(t1.l[k], E1[k], O1[k + 1]);
  addc(t1.l[N - 1], E1[N - 1], 0);
  add_cc(t2.l[0], E2[0], O2[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(t2.l[k], E2[k], O2[k + 1]);
  addc(t2.l[N - 1], E2[N - 1], 0);
  fe_reduce_once<C>(t1); fe_reduce_once<C>(t2);
  r1 = t1; r2 = t2;
}

// ---- separated product + Montgomery reduction (used by the dedicated squaring) ---------------------------------
// One reduction row on the interleaved accumulators (see mont_row): m = low limb * np, add m*q, and take in the next
// limb `tn` of the double-width product at the top.  `c` collects the (rare) overflow bit of the frame's top limb.
template <class C, bool FIRST>
B200_DI void red_row(uint32_t (&E)[C::N], uint32_t (&O)[C::N], uint32_t tn, uint32_t c2, uint32_t& c) {
  constexpr int N = C::N;
  uint32_t m;
  if (FIRST) {
    m = E[0] * C::NP;
    mad_lo_cc(O[0], C::q(1), m, O[0]);
    madc_hi_cc(O[1], C::q(1), m, O[1]);
#pragma unroll
    for (int j = 3; j < N; j += 2) { madc_lo_cc(O[j - 1], C::q(j), m, O[j - 1]); madc_hi_cc(O[j], C::q(j), m, O[j]); }
    addc(c, 0, 0);
  } else {
    add_cc(E[0], E[0], O[1]);
    m = E[0] * C::NP;
#pragma unroll
    for (int j = 1; j < N - 1; j += 2) { madc_lo_cc(O[j - 1], C::q(j), m, O[j + 1]); madc_hi_cc(O[j], C::q(j), m, O[j + 2]); }
    madc_lo_cc(O[N - 2], C::q(N - 1), m, 0);
    madc_hi_cc(O[N - 1], C::q(N - 1), m, tn);
    addc(c, c2, 0);
  }
  mad_lo_cc(E[0], C::q(0), m, E[0]);
  madc_hi_cc(E[1], C::q(0), m, E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) { madc_lo_cc(E[j], C::q(j), m, E[j]); madc_hi_cc(E[j + 1], C::q(j), m, E[j + 1]); }
  addc_cc(O[N - 1], O[N - 1], 0);
  addc(c, c, 0);
}
// r = T / R mod q for a 2N-limb T < q * 2^(32N); result fully reduced.  N*N + N limb products.
template <class C> B200_DI void mont_reduce(Fe<C::N>& r, const uint32_t (&T)[2 * C::N]) {
  constexpr int N = C::N;
  uint32_t E[N], O[N], c = 0;
#pragma unroll
  for (int k = 0; k < N; k++) { E[k] = T[k]; O[k] = 0; }
  O[N - 1] = T[N];
  red_row<C, true>(E, O, 0, 0, c);
#pragma unroll
  for (int i = 1; i < N; i++) {
    uint32_t tn, c2;
    add_cc(tn, T[N + i], c);
    addc(c2, 0, 0);
    if (i & 1) red_row<C, false>(O, E, tn, c2, c); else red_row<C, false>(E, O, tn, c2, c);
  }
  Fe<N> t;
  add_cc(t.l[0], E[0], O[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(t.l[k], E[k], O[k + 1]);
  addc(t.l[N - 1], E[N - 1], 0);
  fe_reduce_once<C>(t);
  r = t;
}
// T = a * a (2N limbs): off-diagonal products a_i*a_j (i < j) in two interleaved accumulators (even / odd position), each row
// one carry chain per parity; then doubled, then the N squares a_i^2 are added in one chain.  N(N-1)/2 + N limb products.
template <class C> B200_DI void sqr_product(uint32_t (&T)[2 * C::N], const uint32_t (&a)[C::N]) {
  constexpr int N = C::N;
  uint32_t Ev[2 * N], Od[2 * N];
#pragma unroll
  for (int k = 0; k < 2 * N; k++) { Ev[k] = 0; Od[k] = 0; }
#pragma unroll
  for (int i = 0; i < N - 1; i++) {
    {   // odd positions: j = i+1, i+3, ...   product at position i+j held in Od[i+j-1], Od[i+j]
      mad_lo_cc(Od[2 * i], a[i], a[i + 1], Od[2 * i]);
      madc_hi_cc(Od[2 * i + 1], a[i], a[i + 1], Od[2 * i + 1]);
      int last = i + 1;
#pragma unroll
      for (int j = i + 3; j < N; j += 2) { madc_lo_cc(Od[i + j - 1], a[i], a[j], Od[i + j - 1]); madc_hi_cc(Od[i + j], a[i], a[j], Od[i + j]); last = j; }
      addc(Od[i + last + 1], Od[i + last + 1], 0);
    }
    if (i + 2 < N) {   // even positions: j = i+2, i+4, ...   product held in Ev[i+j], Ev[i+j+1]
      mad_lo_cc(Ev[2 * i + 2], a[i], a[i + 2], Ev[2 * i + 2]);
      madc_hi_cc(Ev[2 * i + 3], a[i], a[i + 2], Ev[2 * i + 3]);
      int last = i + 2;
#pragma unroll
      for (int j = i + 4; j < N; j += 2) { madc_lo_cc(Ev[i + j], a[i], a[j], Ev[i + j]); madc_hi_cc(Ev[i + j + 1], a[i], a[j], Ev[i + j + 1]); last = j; }
      addc(Ev[i + last + 2], Ev[i + last + 2], 0);
    }
  }
  T[0] = Ev[0];
  add_cc(T[1], Ev[1], Od[0]);
#pragma unroll
  for (int p = 2; p < 2 * N - 1; p++) addc_cc(T[p], Ev[p], Od[p - 1]);
  addc(T[2 * N - 1], Ev[2 * N - 1], Od[2 * N - 2]);
#pragma unroll
  for (int p = 2 * N - 1; p > 0; p--) T[p] = __funnelshift_l(T[p - 1], T[p], 1);
  T[0] <<= 1;
  mad_lo_cc(T[0], a[0], a[0], T[0]);
  madc_hi_cc(T[1], a[0], a[0], T[1]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) { madc_lo_cc(T[2 * i], a[i], a[i], T[2 * i]); madc_hi_cc(T[2 * i + 1], a[i], a[i], T[2 * i + 1]); }
  madc_lo_cc(T[2 * N - 2], a[N - 1], a[N - 1], T[2 * N - 2]);
  madc_hi(T[2 * N - 1], a[N - 1], a[N - 1], T[2 * N - 1]);
}
#if defined(B200_EXPERIMENTS)      // measured alternative (no gain): kept out of the shipped library
// H x H limb product (2H limbs) in the same two-accumulator carry-chain form: every row x*y_i is two chains of H/2 wide
// multiply-adds (even j, odd j); the word after a chain's end has not been touched by earlier rows, so its carry just lands.
template <int H> B200_DI void half_product(uint32_t (&T)[2 * H], const uint32_t* x, const uint32_t* y) {
  uint32_t Ev[2 * H + 2], Od[2 * H + 2];
#pragma unroll
  for (int k = 0; k < 2 * H + 2; k++) { Ev[k] = 0; Od[k] = 0; }
#pragma unroll
  for (int i = 0; i < H; i++) {
#pragma unroll
    for (int par = 0; par < 2; par++) {
      const int p0 = i + par;
      if ((p0 & 1) == 0) { mad_lo_cc(Ev[p0], x[par], y[i], Ev[p0]); madc_hi_cc(Ev[p0 + 1], x[par], y[i], Ev[p0 + 1]); }
      else { mad_lo_cc(Od[p0 - 1], x[par], y[i], Od[p0 - 1]); madc_hi_cc(Od[p0], x[par], y[i], Od[p0]); }
      int last = par;
#pragma unroll
      for (int j = par + 2; j < H; j += 2) {
        const int p = i + j;
        if ((p & 1) == 0) { madc_lo_cc(Ev[p], x[j], y[i], Ev[p]); madc_hi_cc(Ev[p + 1], x[j], y[i], Ev[p + 1]); }
        else { madc_lo_cc(Od[p - 1], x[j], y[i], Od[p - 1]); madc_hi_cc(Od[p], x[j], y[i], Od[p]); }
        last = j;
      }
      const int pl = i + last;
      if ((pl & 1) == 0) addc(Ev[pl + 2], Ev[pl + 2], 0); else addc(Od[pl + 1], Od[pl + 1], 0);
    }
  }
  T[0] = Ev[0];
  add_cc(T[1], Ev[1], Od[0]);
#pragma unroll
  for (int p = 2; p < 2 * H - 1; p++) addc_cc(T[p], Ev[p], Od[p - 1]);
  addc(T[2 * H - 1], Ev[2 * H - 1], Od[2 * H - 2]);
}
// One level of Karatsuba on the a*b half of the Montgomery product: three (N/2)x(N/2) products instead of NxN, i.e.
// 3N^2/4 + N^2 + N limb products (264 for BLS12-381, 120 for BN254) instead of 2N^2 + N, paid for with ~100 extra
// additions on the ALU pipe, which has slack while IMAD.WIDE (32 lanes/clk/SM) is the bound.
template <class C> B200_DI void fe_mul_karatsuba(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  constexpr int N = C::N, H = N / 2;
  uint32_t T[2 * N], zm[2 * H + 1], sa[H], sb[H], ca, cb;
  {
    uint32_t z0[2 * H], z2[2 * H];
    half_product<H>(z0, a.l, b.l);
    half_product<H>(z2, a.l + H, b.l + H);
    add_cc(sa[0], a.l[0], a.l[H]);
#pragma unroll
    for (int k = 1; k < H; k++) addc_cc(sa[k], a.l[k], a.l[H + k]);
    addc(ca, 0, 0);
    add_cc(sb[0], b.l[0], b.l[H]);
#pragma unroll
    for (int k = 1; k < H; k++) addc_cc(sb[k], b.l[k], b.l[H + k]);
    addc(cb, 0, 0);
    {
      uint32_t zz[2 * H];
      half_product<H>(zz, sa, sb);
#pragma unroll
      for (int k = 0; k < 2 * H; k++) zm[k] = zz[k];
      zm[2 * H] = 0;
    }
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    add_cc(zm[H], zm[H], sb[0] & ma);
#pragma unroll
    for (int k = 1; k < H; k++) addc_cc(zm[H + k], zm[H + k], sb[k] & ma);
    addc(zm[2 * H], zm[2 * H], 0);
    add_cc(zm[H], zm[H], sa[0] & mb);
#pragma unroll
    for (int k = 1; k < H; k++) addc_cc(zm[H + k], zm[H + k], sa[k] & mb);
    addc(zm[2 * H], zm[2 * H], ca & cb);
    sub_cc(zm[0], zm[0], z0[0]);
#pragma unroll
    for (int k = 1; k < 2 * H; k++) subc_cc(zm[k], zm[k], z0[k]);
    subc(zm[2 * H], zm[2 * H], 0);
    sub_cc(zm[0], zm[0], z2[0]);
#pragma unroll
    for (int k = 1; k < 2 * H; k++) subc_cc(zm[k], zm[k], z2[k]);
    subc(zm[2 * H], zm[2 * H], 0);
#pragma unroll
    for (int k = 0; k < 2 * H; k++) { T[k] = z0[k]; T[2 * H + k] = z2[k]; }
  }
  add_cc(T[H], T[H], zm[0]);
#pragma unroll
  for (int k = 1; k < 2 * H + 1; k++) addc_cc(T[H + k], T[H + k], zm[k]);
#pragma unroll
  for (int k = 3 * H + 1; k < 2 * N - 1; k++) addc_cc(T[k], T[k], 0);
  addc(T[2 * N - 1], T[2 * N - 1], 0);
  mont_reduce<C>(r, T);
}
#endif
template <class C> B200_DI void fe_mul_p(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
#if defined(B200_EXPERIMENTS) && defined(B200_KARATSUBA)
  fe_mul_karatsuba<C>(r, a, b);
#else
  fe_mul_cios<C>(r, a, b);
#endif
}

// f1m_square (build_f1m.js:779-1076): dedicated squaring, N(N-1)/2 + N + N*N + N limb products (234 for BLS12-381, 108 for BN254)
// instead of the 2N*N + N of a general multiplication.
template <class C> B200_DI void fe_sqr_p(Fe<C::N>& r, const Fe<C::N>& a) {
#if defined(B200_SQR_AS_MUL)
  fe_mul<C>(r, a, a);
#else
  uint32_t T[2 * C::N];
  sqr_product<C>(T, a.l);
  mont_reduce<C>(r, T);
#endif
}

// to / from Montgomery (build_f1m.js:1089,1098)
template <class C> B200_DI void fe_to_mont_p(Fe<C::N>& r, const Fe<C::N>& a) {
  Fe<C::N> k;
#pragma unroll
  for (int i = 0; i < C::N; i++) k.l[i] = C::r2(i);
  fe_mul<C>(r, a, k);
}
template <class C> B200_DI void fe_from_mont_p(Fe<C::N>& r, const Fe<C::N>& a) {
  Fe<C::N> k;
#pragma unroll
  for (int i = 0; i < C::N; i++) k.l[i] = (i == 0);
  fe_mul<C>(r, a, k);
}

// f1m_inverse (build_f1m.js:1112-1122) by Fermat: a^(q-2).  Montgomery in, Montgomery out; inv(0) = 0.
// Left-to-right square-and-multiply over the constant exponent; the loop is NOT unrolled (code size).
template <class C> __device__ __noinline__ void fe_inv_p(Fe<C::N>& r, const Fe<C::N>& a) {
  constexpr int N = C::N;
  uint32_t e[N];
  // e = q - 2 (q is odd and its low limb is >= 2 for both fields)
#pragma unroll
  for (int i = 0; i < N; i++) e[i] = C::q(i);
  e[0] -= 2;
  Fe<N> acc; fe_set_one<C>(acc);
  for (int i = C::QBITS - 1; i >= 0; i--) {
    fe_sqr<C>(acc, acc);
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < N; k++) w = (k == (i >> 5)) ? e[k] : w;
    if ((w >> (i & 31)) & 1) fe_mul<C>(acc, acc, a);
  }
  r = acc;
}

// f1m_inverse for the latency-critical call sites (the root of every batch inversion, k_normalize, the codecs): Pornin's optimised
// binary GCD (bingcd.h: 31 shift/subtract steps at a time on 64-bit approximations, ~5x fewer instructions than the plain binary
// Euclid below and an order of magnitude fewer than Fermat).  The reference also uses an extended Euclid here
// (build_int.js:922-1064).  Variable time; inv(0) = 0.  Montgomery in, Montgomery out.
template <class C> __device__ __noinline__ void fe_inv_fast_p(Fe<C::N>& r, const Fe<C::N>& a) {
  constexpr int N = C::N;
  uint32_t x[N];
  bingcd_inverse<C>(x, a.l);
  // x = (aR)^-1 as a plain residue; a^-1 R = x * R^2 = montmul(x, R^3)
  Fe<N> t, k;
#pragma unroll
  for (int i = 0; i < N; i++) { t.l[i] = x[i]; k.l[i] = C::r3(i); }
  fe_mul<C>(r, t, k);
}

#if defined(B200_EXPERIMENTS)
// f1m_inverse by the binary extended Euclidean algorithm (right-shift form): ~2*log2(q) shift/subtract
// steps on N-limb integers instead of ~1.5*log2(q) field multiplications -- an order of magnitude fewer
// instructions than Fermat, which matters because the batch inversion's root is a single serial chain.
// The reference also uses an extended Euclid here (build_int.js:922-1064).  Variable time; inv(0) = 0.
template <class C> B200_DI void limbs_shr1(uint32_t (&a)[C::N], uint32_t top) {
#pragma unroll
  for (int i = 0; i < C::N - 1; i++) a[i] = __funnelshift_r(a[i], a[i + 1], 1);
  a[C::N - 1] = (a[C::N - 1] >> 1) | (top << 31);
}
template <class C> B200_DI void limbs_halve_mod(uint32_t (&x)[C::N]) {     // x <- x / 2 mod q
  constexpr int N = C::N;
  uint32_t carry = 0;
  if (x[0] & 1) {
    add_cc(x[0], x[0], C::q(0));
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(x[i], x[i], C::q(i));
    addc(carry, 0, 0);
  }
  limbs_shr1<C>(x, carry);
}
template <class C> B200_DI bool limbs_is_one(const uint32_t (&a)[C::N]) {
  uint32_t o = a[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < C::N; i++) o |= a[i];
  return o == 0;
}
template <class C> __device__ __noinline__ void fe_inv_euclid_p(Fe<C::N>& r, const Fe<C::N>& a) {
  constexpr int N = C::N;
  if (fe_is_zero<C>(a)) { fe_set_zero<C>(r); return; }
  uint32_t u[N], v[N];
  Fe<N> x1, x2;
#pragma unroll
  for (int i = 0; i < N; i++) { u[i] = a.l[i]; v[i] = C::q(i); x1.l[i] = (i == 0); x2.l[i] = 0; }
  for (;;) {
    while (!(u[0] & 1)) { limbs_shr1<C>(u, 0); limbs_halve_mod<C>(x1.l); }
    if (limbs_is_one<C>(u)) { x2 = x1; break; }
    while (!(v[0] & 1)) { limbs_shr1<C>(v, 0); limbs_halve_mod<C>(x2.l); }
    if (limbs_is_one<C>(v)) break;
    // u, v odd and different: subtract the smaller from the larger
    uint32_t t[N], borrow;
    sub_cc(t[0], u[0], v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], u[i], v[i]);
    subc(borrow, 0, 0);
    if (!borrow) {
#pragma unroll
      for (int i = 0; i < N; i++) u[i] = t[i];
      fe_sub<C>(x1, x1, x2);
    } else {
      sub_cc(v[0], v[0], u[0]);
#pragma unroll
      for (int i = 1; i < N - 1; i++) subc_cc(v[i], v[i], u[i]);
      subc(v[N - 1], v[N - 1], u[N - 1]);
      fe_sub<C>(x2, x2, x1);
    }
  }
  // x2 = (aR)^-1 as a plain residue; a^-1 R = x2 * R^2 = montmul(x2, R^3)
  Fe<N> k;
#pragma unroll
  for (int i = 0; i < N; i++) k.l[i] = C::r3(i);
  fe_mul<C>(r, x2, k);
}

#endif  // B200_EXPERIMENTS

// ------------------------------------------------------------------ Fq2 arithmetic on the halves + the dispatching entry points
template <class C> B200_DI void fq2_get(Fe<C::Base::N>& a0, Fe<C::Base::N>& a1, const Fe<C::N>& a) {
#pragma unroll
  for (int i = 0; i < C::Base::N; i++) { a0.l[i] = a.l[i]; a1.l[i] = a.l[C::Base::N + i]; }
}
template <class C> B200_DI void fq2_put(Fe<C::N>& r, const Fe<C::Base::N>& r0, const Fe<C::Base::N>& r1) {
#pragma unroll
  for (int i = 0; i < C::Base::N; i++) { r.l[i] = r0.l[i]; r.l[C::Base::N + i] = r1.l[i]; }
}
template <class C> B200_DI void fe_add(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  if constexpr (C::EXT == 2) { using B = typename C::Base; Fe<B::N> a0, a1, b0, b1; fq2_get<C>(a0, a1, a); fq2_get<C>(b0, b1, b); fe_add_p<B>(a0, a0, b0); fe_add_p<B>(a1, a1, b1); fq2_put<C>(r, a0, a1); }
  else fe_add_p<C>(r, a, b);
}
template <class C> B200_DI void fe_sub(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  if constexpr (C::EXT == 2) { using B = typename C::Base; Fe<B::N> a0, a1, b0, b1; fq2_get<C>(a0, a1, a); fq2_get<C>(b0, b1, b); fe_sub_p<B>(a0, a0, b0); fe_sub_p<B>(a1, a1, b1); fq2_put<C>(r, a0, a1); }
  else fe_sub_p<C>(r, a, b);
}
template <class C> B200_DI void fe_neg(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) { using B = typename C::Base; Fe<B::N> a0, a1; fq2_get<C>(a0, a1, a); fe_neg_p<B>(a0, a0); fe_neg_p<B>(a1, a1); fq2_put<C>(r, a0, a1); }
  else fe_neg_p<C>(r, a);
}
// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u: 3 base multiplications      (f2m_mul, build_f2m.js:152-194)
template <class C> B200_DI void fe_mul(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) {
  if constexpr (C::EXT == 2) {
    using B = typename C::Base; Fe<B::N> a0, a1, b0, b1, v0, v1, sa, sb;
    fq2_get<C>(a0, a1, a); fq2_get<C>(b0, b1, b);
    fe_add_p<B>(sa, a0, a1); fe_add_p<B>(sb, b0, b1);
    fe_mul_p<B>(v0, a0, b0); fe_mul_p<B>(v1, a1, b1); fe_mul_p<B>(sa, sa, sb);
    fe_sub_p<B>(sa, sa, v0); fe_sub_p<B>(sa, sa, v1); fe_sub_p<B>(v0, v0, v1);
    fq2_put<C>(r, v0, sa);
  } else fe_mul_p<C>(r, a, b);
}
// (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u: 2 base multiplications                                         (f2m_square, build_f2m.js:290-330)
template <class C> B200_DI void fe_sqr(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) {
    using B = typename C::Base; Fe<B::N> a0, a1, s, d, p;
    fq2_get<C>(a0, a1, a);
    fe_add_p<B>(s, a0, a1); fe_sub_p<B>(d, a0, a1); fe_mul_p<B>(p, a0, a1);
    fe_mul_p<B>(s, s, d); fe_add_p<B>(p, p, p);
    fq2_put<C>(r, s, p);
  } else fe_sqr_p<C>(r, a);
}
template <class C> B200_DI void fe_to_mont(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) { using B = typename C::Base; Fe<B::N> a0, a1; fq2_get<C>(a0, a1, a); fe_to_mont_p<B>(a0, a0); fe_to_mont_p<B>(a1, a1); fq2_put<C>(r, a0, a1); }
  else fe_to_mont_p<C>(r, a);
}
template <class C> B200_DI void fe_from_mont(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) { using B = typename C::Base; Fe<B::N> a0, a1; fq2_get<C>(a0, a1, a); fe_from_mont_p<B>(a0, a0); fe_from_mont_p<B>(a1, a1); fq2_put<C>(r, a0, a1); }
  else fe_from_mont_p<C>(r, a);
}
// 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + a1^2)                                                                  (f2m_inverse, build_f2m.js:402-440)
template <class C, bool FAST> B200_DI void fq2_inv(Fe<C::N>& r, const Fe<C::N>& a) {
  using B = typename C::Base; Fe<B::N> a0, a1, t0, t1;
  fq2_get<C>(a0, a1, a);
  fe_sqr_p<B>(t0, a0); fe_sqr_p<B>(t1, a1); fe_add_p<B>(t0, t0, t1);
  if (FAST) fe_inv_fast_p<B>(t1, t0); else fe_inv_p<B>(t1, t0);
  fe_mul_p<B>(a0, a0, t1); fe_mul_p<B>(a1, a1, t1); fe_neg_p<B>(a1, a1);
  fq2_put<C>(r, a0, a1);
}
template <class C> B200_DI void fe_inv(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) fq2_inv<C, false>(r, a); else fe_inv_p<C>(r, a);
}
template <class C> B200_DI void fe_inv_fast(Fe<C::N>& r, const Fe<C::N>& a) {
  if constexpr (C::EXT == 2) fq2_inv<C, true>(r, a); else fe_inv_fast_p<C>(r, a);
}

// Multiplication / squaring for the XYZZ formulas (ec.cuh).  Over Fq2 they are CALLS: an XYZZ addition is 14 Fq2 = 42 Fq
// multiplications, and inlining them all pushes the fold / finish / table kernels to 255 registers plus a kilobyte of spill
// (and the compile of the two G2 instantiations to ten minutes).  The batch-affine tree kernels keep the inlined forms.
template <class C> __device__ __noinline__ void fe_mul_call(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) { Fe<C::N> t; fe_mul<C>(t, a, b); r = t; }
template <class C> __device__ __noinline__ void fe_sqr_call(Fe<C::N>& r, const Fe<C::N>& a) { Fe<C::N> t; fe_sqr<C>(t, a); r = t; }
template <class C> B200_DI void fe_mul_x(Fe<C::N>& r, const Fe<C::N>& a, const Fe<C::N>& b) { if constexpr (C::EXT == 2) fe_mul_call<C>(r, a, b); else fe_mul<C>(r, a, b); }
template <class C> B200_DI void fe_sqr_x(Fe<C::N>& r, const Fe<C::N>& a) { if constexpr (C::EXT == 2) fe_sqr_call<C>(r, a); else fe_sqr<C>(r, a); }

// ------------------------------------------------------------------ global-memory access helpers
// Elements are 16-byte aligned in every buffer the engine owns; 128-bit vector loads/stores.
template <class C> B200_DI void fe_load(Fe<C::N>& r, const void* p) {
  const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < C::N / 4; i++) { uint4 v = __ldg(s + i); r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w; }
}
template <class C> B200_DI void fe_load_cg(Fe<C::N>& r, const void* p) {      // coherent (written earlier in the same launch sequence)
  const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < C::N / 4; i++) { uint4 v = s[i]; r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w; }
}
template <class C> B200_DI void fe_load_l2(Fe<C::N>& r, const void* p) {      // ld.global.cg: served by L2, never by a (possibly stale) L1 line -- values another CTA wrote during this launch
  const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < C::N / 4; i++) { uint4 v = __ldcg(s + i); r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w; }
}
template <class C> B200_DI void fe_store(void* p, const Fe<C::N>& a) {
  uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < C::N / 4; i++) d[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
}

}  // namespace b200
