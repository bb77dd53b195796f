// ec.cuh -- short-Weierstrass a=0 group law for G1 (BLS12-381, BN254) on top of fp.cuh.
//
// GPU counterpart of wasmcurves/src/build_curve_jacobian_a0.js:
//   g1m_add :541-658, g1m_addMixed :661-761, g1m_double :291-359, g1m_zero :124-150,
//   g1m_isZeroAffine :55-77 (affine infinity = (0,0)).
// The reference accumulates buckets in Jacobian coordinates; here the running sums use the
// extended-Jacobian "XYZZ" form (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), which makes the mixed add
// 8M+2S with no field inversions, and the bulk of bucket accumulation uses affine + affine with
// a shared (batched) inversion, as the reference's opt path does (build_multiexp_opt.js:1016-1245).
// All functions are complete: infinity, P+P and P+(-P) are handled (the reference's batch-affine
// formulas are not -- SURVEY.md 8a defects 1-3).
#pragma once
#include "fp.cuh"

namespace b200 {

template <class C> struct Affine { Fe<C::N> x, y; };            // infinity: x == y == 0
template <class C> struct XYZZ { Fe<C::N> x, y, zz, zzz; };     // infinity: zz == 0

template <class C> B200_DI bool affine_is_inf(const Affine<C>& p) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < C::N; i++) o |= p.x.l[i] | p.y.l[i];
  return o == 0;
}
template <class C> B200_DI void affine_set_inf(Affine<C>& p) { fe_set_zero<C>(p.x); fe_set_zero<C>(p.y); }
template <class C> B200_DI bool xyzz_is_inf(const XYZZ<C>& p) { return fe_is_zero<C>(p.zz); }
template <class C> B200_DI void xyzz_set_inf(XYZZ<C>& p) {
  fe_set_zero<C>(p.x); fe_set_one<C>(p.y); fe_set_zero<C>(p.zz); fe_set_zero<C>(p.zzz);
}
template <class C> B200_DI void xyzz_from_affine(XYZZ<C>& r, const Affine<C>& p) {
  if (affine_is_inf<C>(p)) { xyzz_set_inf<C>(r); return; }
  r.x = p.x; r.y = p.y; fe_set_one<C>(r.zz); fe_set_one<C>(r.zzz);
}

template <class C> B200_DI void affine_load(Affine<C>& p, const void* base, uint64_t idx) {
  const char* s = reinterpret_cast<const char*>(base) + idx * (uint64_t)(8 * C::N);
  fe_load<C>(p.x, s); fe_load<C>(p.y, s + 4 * C::N);
}
template <class C> B200_DI void affine_load_cg(Affine<C>& p, const void* base, uint64_t idx) {
  const char* s = reinterpret_cast<const char*>(base) + idx * (uint64_t)(8 * C::N);
  fe_load_cg<C>(p.x, s); fe_load_cg<C>(p.y, s + 4 * C::N);
}
template <class C> B200_DI void affine_store(void* base, uint64_t idx, const Affine<C>& p) {
  char* d = reinterpret_cast<char*>(base) + idx * (uint64_t)(8 * C::N);
  fe_store<C>(d, p.x); fe_store<C>(d + 4 * C::N, p.y);
}
template <class C> B200_DI void xyzz_load(XYZZ<C>& p, const void* base, uint64_t idx) {
  const char* s = reinterpret_cast<const char*>(base) + idx * (uint64_t)(16 * C::N);
  fe_load_cg<C>(p.x, s); fe_load_cg<C>(p.y, s + 4 * C::N); fe_load_cg<C>(p.zz, s + 8 * C::N); fe_load_cg<C>(p.zzz, s + 12 * C::N);
}
template <class C> B200_DI void xyzz_store(void* base, uint64_t idx, const XYZZ<C>& p) {
  char* d = reinterpret_cast<char*>(base) + idx * (uint64_t)(16 * C::N);
  fe_store<C>(d, p.x); fe_store<C>(d + 4 * C::N, p.y); fe_store<C>(d + 8 * C::N, p.zz); fe_store<C>(d + 12 * C::N, p.zzz);
}

// 2*P for affine P != inf  (mdbl-2008-s-1; y != 0 on these curves: no 2-torsion)
template <class C> __device__ __noinline__ void xyzz_dbl_affine(XYZZ<C>& r, const Affine<C>& p) {
  Fe<C::N> U, V, W, S, M, t;
  fe_dbl<C>(U, p.y); fe_sqr_x<C>(V, U); fe_mul_x<C>(W, U, V); fe_mul_x<C>(S, p.x, V);
  fe_sqr_x<C>(t, p.x); fe_dbl<C>(M, t); fe_add<C>(M, M, t);
  fe_sqr_x<C>(r.x, M); fe_sub<C>(r.x, r.x, S); fe_sub<C>(r.x, r.x, S);
  fe_sub<C>(t, S, r.x); fe_mul_x<C>(t, M, t); fe_mul_x<C>(U, W, p.y); fe_sub<C>(r.y, t, U);
  r.zz = V; r.zzz = W;
}

// 2*P in XYZZ (dbl-2008-s-1, a = 0).  g1m_double, build_curve_jacobian_a0.js:291-359
template <class C> __device__ __noinline__ void xyzz_dbl(XYZZ<C>& r, const XYZZ<C>& p) {
  if (xyzz_is_inf<C>(p)) { r = p; return; }
  Fe<C::N> U, V, W, S, M, t, X3;
  fe_dbl<C>(U, p.y); fe_sqr_x<C>(V, U); fe_mul_x<C>(W, U, V); fe_mul_x<C>(S, p.x, V);
  fe_sqr_x<C>(t, p.x); fe_dbl<C>(M, t); fe_add<C>(M, M, t);
  fe_sqr_x<C>(X3, M); fe_sub<C>(X3, X3, S); fe_sub<C>(X3, X3, S);
  fe_sub<C>(t, S, X3); fe_mul_x<C>(t, M, t); fe_mul_x<C>(U, W, p.y);
  r.x = X3; fe_sub<C>(r.y, t, U);
  fe_mul_x<C>(r.zz, V, p.zz); fe_mul_x<C>(r.zzz, W, p.zzz);
}

// acc += P (affine).  g1m_addMixed, build_curve_jacobian_a0.js:661-761 (madd-2008-s in XYZZ)
template <class C> B200_DI void xyzz_madd(XYZZ<C>& acc, const Affine<C>& p) {
  if (affine_is_inf<C>(p)) return;
  if (xyzz_is_inf<C>(acc)) { acc.x = p.x; acc.y = p.y; fe_set_one<C>(acc.zz); fe_set_one<C>(acc.zzz); return; }
  Fe<C::N> U2, S2, P, R, PP, PPP, Q, t;
  fe_mul_x<C>(U2, p.x, acc.zz); fe_mul_x<C>(S2, p.y, acc.zzz);
  fe_sub<C>(P, U2, acc.x); fe_sub<C>(R, S2, acc.y);
  if (fe_is_zero<C>(P)) {
    if (fe_is_zero<C>(R)) xyzz_dbl_affine<C>(acc, p); else xyzz_set_inf<C>(acc);
    return;
  }
  fe_sqr_x<C>(PP, P); fe_mul_x<C>(PPP, P, PP); fe_mul_x<C>(Q, acc.x, PP);
  fe_sqr_x<C>(t, R); fe_sub<C>(t, t, PPP); fe_sub<C>(t, t, Q); fe_sub<C>(acc.x, t, Q);
  fe_sub<C>(t, Q, acc.x); fe_mul_x<C>(t, R, t); fe_mul_x<C>(Q, acc.y, PPP); fe_sub<C>(acc.y, t, Q);
  fe_mul_x<C>(acc.zz, acc.zz, PP); fe_mul_x<C>(acc.zzz, acc.zzz, PPP);
}

// acc += Q (XYZZ).  g1m_add, build_curve_jacobian_a0.js:541-658 (add-2008-s)
template <class C> B200_DI void xyzz_add(XYZZ<C>& acc, const XYZZ<C>& q) {
  if (xyzz_is_inf<C>(q)) return;
  if (xyzz_is_inf<C>(acc)) { acc = q; return; }
  Fe<C::N> U1, U2, S1, S2, P, R, PP, PPP, Q, t;
  fe_mul_x<C>(U1, acc.x, q.zz); fe_mul_x<C>(U2, q.x, acc.zz);
  fe_mul_x<C>(S1, acc.y, q.zzz); fe_mul_x<C>(S2, q.y, acc.zzz);
  fe_sub<C>(P, U2, U1); fe_sub<C>(R, S2, S1);
  if (fe_is_zero<C>(P)) {
    if (fe_is_zero<C>(R)) { XYZZ<C> d; xyzz_dbl<C>(d, q); acc = d; } else xyzz_set_inf<C>(acc);
    return;
  }
  fe_sqr_x<C>(PP, P); fe_mul_x<C>(PPP, P, PP); fe_mul_x<C>(Q, U1, PP);
  fe_sqr_x<C>(t, R); fe_sub<C>(t, t, PPP); fe_sub<C>(t, t, Q); fe_sub<C>(acc.x, t, Q);
  fe_sub<C>(t, Q, acc.x); fe_mul_x<C>(t, R, t); fe_mul_x<C>(Q, S1, PPP); fe_sub<C>(acc.y, t, Q);
  fe_mul_x<C>(t, acc.zz, q.zz); fe_mul_x<C>(acc.zz, t, PP);
  fe_mul_x<C>(t, acc.zzz, q.zzz); fe_mul_x<C>(acc.zzz, t, PPP);
}

// (Measured and dropped, round 2: the same addition with its fourteen multiplications issued as seven interleaved PAIRS (fe_mul2) in the latency-bound
// fold tail -- 0.445 vs 0.314 ms at 2^20: the SM issues in order, so a second carry chain in the same thread adds instructions without hiding latency.)

// XYZZ -> Jacobian (X', Y', Z') with x = X'/Z'^2, y = Y'/Z'^3, no inversion:
// Z' = ZZ*ZZZ, X' = X*ZZ*ZZZ^2, Y' = Y*ZZ^3*ZZZ^2.  Infinity -> canonical zero (0, R mod q, 0) = g1m_zero :124-150.
template <class C> B200_DI void xyzz_to_jacobian(Fe<C::N>& X, Fe<C::N>& Y, Fe<C::N>& Z, const XYZZ<C>& p) {
  if (xyzz_is_inf<C>(p)) { fe_set_zero<C>(X); fe_set_one<C>(Y); fe_set_zero<C>(Z); return; }
  Fe<C::N> t, u;
  fe_mul_x<C>(Z, p.zz, p.zzz);            // Z'
  fe_mul_x<C>(t, Z, p.zzz);               // ZZ*ZZZ^2
  fe_mul_x<C>(X, p.x, t);
  fe_sqr_x<C>(u, p.zz); fe_mul_x<C>(t, t, u);   // ZZ^3*ZZZ^2
  fe_mul_x<C>(Y, p.y, t);
}

// Decision + denominator for one affine + affine addition that shares a batched inversion
// (pass 1 of build_multiexp_opt.js:1090-1126).  kind: 0 = generic (d = x2 - x1), 1 = doubling (d = 2*y1),
// 2 = result is p2 (p1 = inf), 3 = result is p1 (p2 = inf), 4 = result is infinity (p1 = -p2).
// For kinds 2..4 the denominator is 1 so the running product is unaffected.
template <class C> B200_DI int affine_add_denominator(Fe<C::N>& d, const Affine<C>& p1, const Affine<C>& p2) {
  if (affine_is_inf<C>(p1)) { fe_set_one<C>(d); return 2; }
  if (affine_is_inf<C>(p2)) { fe_set_one<C>(d); return 3; }
  fe_sub<C>(d, p2.x, p1.x);
  if (fe_is_zero<C>(d)) {
    // P + P with y == 0 (a 2-torsion point: impossible in the prime-order groups, possible for unchecked input) doubles to infinity;
    // it must stay out of the shared product -- d = 2y = 0 would zero the whole batch inversion
    if (fe_eq<C>(p1.y, p2.y) && !fe_is_zero<C>(p1.y)) { fe_dbl<C>(d, p1.y); return 1; }
    fe_set_one<C>(d); return 4;
  }
  return 0;
}
// Pass 2 (build_multiexp_opt.js:1178-1239): given dinv = 1/d, finish the addition.
template <class C> B200_DI void affine_add_finish(Affine<C>& r, const Affine<C>& p1, const Affine<C>& p2, const Fe<C::N>& dinv, int kind) {
  if (kind == 2) { r = p2; return; }
  if (kind == 3) { r = p1; return; }
  if (kind == 4) { affine_set_inf<C>(r); return; }
  Fe<C::N> lam, t, x3;
  if (kind == 0) { fe_sub<C>(t, p2.y, p1.y); }
  else { fe_sqr<C>(lam, p1.x); fe_dbl<C>(t, lam); fe_add<C>(t, t, lam); }
  fe_mul<C>(lam, t, dinv);
  fe_sqr<C>(x3, lam); fe_sub<C>(x3, x3, p1.x); fe_sub<C>(x3, x3, p2.x);
  fe_sub<C>(t, p1.x, x3); fe_mul<C>(t, lam, t); fe_sub<C>(r.y, t, p1.y);
  r.x = x3;
}

}  // namespace b200
