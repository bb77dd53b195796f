// host_ec.h -- the serial tail of the MSM on one host core.
//
// After the GPU has folded every window's bucket array, what remains is
//     result = sum_w 2^(c*w) * ( T_w[0] + sum_j 2^j * T_w[2^j] )
// (accumulateAcrossChunks, wasmcurves/src/build_multiexp_opt.js:1710-1746; multiexp Horner loop,
// wasmcurves/src/build_multiexp.js:319-369): ONE dependent chain of ~nbits point doublings over fewer than 300 points.
// A dependent chain is the worst case for a GPU (measured: 10.7 us per XYZZ doubling on one B200 thread, 2.9 ms for
// 272 of them) and the best case for a CPU core (~0.3 us per doubling), so this step -- and only this step -- runs
// on the host, on the ~50 KB of folded points the device hands back.  It is part of the engine (not a fallback:
// there is no other implementation of this step on the default path) and it is exact integer arithmetic, so the
// result is the same group element the device chain (k_horner, kept for cross-checking) produces.
//
// Field: Montgomery form, R = 2^(64*L), 64-bit limbs via unsigned __int128 (same values as fp.cuh's 32-bit limbs).
#pragma once
#include <vector>
#include <stdint.h>
#include <string.h>

namespace b200host {

typedef unsigned __int128 u128;

// Field descriptors.  W = 64-bit words per element.  Field<L>: prime field Fq.  Field2<L>: Fq2 = Fq[u]/(u^2 + 1), element c0 || c1
// (f2m layout, wasmcurves/src/build_f2m.js; both curves use f1m_neg as the non-residue multiplication) -- the G2 coordinate field.
template <int L> struct Field  { static constexpr int W = L;     uint64_t q[L], one[L], np; };
template <int L> struct Field2 { static constexpr int W = 2 * L; Field<L> b; uint64_t one[2 * L]; };

template <int W> struct Fe { uint64_t l[W]; };

template <int W> static inline bool is_zero(const Fe<W>& a) { uint64_t o = 0; for (int i = 0; i < W; i++) o |= a.l[i]; return o == 0; }
template <int L> static inline bool ge_q(const Field<L>& f, const uint64_t* a) {
  for (int i = L - 1; i >= 0; i--) { if (a[i] > f.q[i]) return true; if (a[i] < f.q[i]) return false; }
  return true;
}
// ---- prime field, on raw word arrays (so that the extension can address its halves)
template <int L> static inline void add_w(const Field<L>& f, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t c = 0;
  for (int i = 0; i < L; i++) { u128 s = (u128)a[i] + b[i] + c; r[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
  if (c || ge_q<L>(f, r)) { uint64_t br = 0; for (int i = 0; i < L; i++) { u128 d = (u128)r[i] - f.q[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } }
}
template <int L> static inline void sub_w(const Field<L>& f, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t br = 0;
  for (int i = 0; i < L; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
  if (br) { uint64_t c = 0; for (int i = 0; i < L; i++) { u128 s = (u128)r[i] + f.q[i] + c; r[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } }
}
// CIOS Montgomery product, "no-carry" form (valid because the top bit of q's top word is clear for both fields, so the
// running value never needs an extra word): per inner step two 64x64->128 products and two additions.  r may alias a or b.
template <int L> static inline __attribute__((always_inline)) void mul_w(const Field<L>& f, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t[L];
  for (int j = 0; j < L; j++) t[j] = 0;
  for (int i = 0; i < L; i++) {
    const uint64_t bi = b[i];
    u128 A = (u128)a[0] * bi + t[0];
    const uint64_t m = (uint64_t)A * f.np;
    u128 Cc = (u128)m * f.q[0] + (uint64_t)A;
    uint64_t ca = (uint64_t)(A >> 64), cc = (uint64_t)(Cc >> 64);
    for (int j = 1; j < L; j++) {
      A = (u128)a[j] * bi + t[j] + ca; ca = (uint64_t)(A >> 64);
      Cc = (u128)m * f.q[j] + (uint64_t)A + cc; cc = (uint64_t)(Cc >> 64);
      t[j - 1] = (uint64_t)Cc;
    }
    t[L - 1] = ca + cc;
  }
  if (ge_q<L>(f, t)) { uint64_t br = 0;
    for (int i = 0; i < L; i++) { u128 d = (u128)t[i] - f.q[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } }
  for (int i = 0; i < L; i++) r[i] = t[i];
}
// ---- the field interface the point formulas use: add / sub / mul / sqr / dbl on Fe<F::W>, overloaded on the descriptor
template <int L> static inline void add(const Field<L>& f, Fe<L>& r, const Fe<L>& a, const Fe<L>& b) { add_w<L>(f, r.l, a.l, b.l); }
template <int L> static inline void sub(const Field<L>& f, Fe<L>& r, const Fe<L>& a, const Fe<L>& b) { sub_w<L>(f, r.l, a.l, b.l); }
template <int L> static inline void mul(const Field<L>& f, Fe<L>& r, const Fe<L>& a, const Fe<L>& b) { mul_w<L>(f, r.l, a.l, b.l); }
template <int L> static inline void sqr(const Field<L>& f, Fe<L>& r, const Fe<L>& a) { mul_w<L>(f, r.l, a.l, a.l); }
template <int L> static inline void add(const Field2<L>& f, Fe<2 * L>& r, const Fe<2 * L>& a, const Fe<2 * L>& b) { add_w<L>(f.b, r.l, a.l, b.l); add_w<L>(f.b, r.l + L, a.l + L, b.l + L); }
template <int L> static inline void sub(const Field2<L>& f, Fe<2 * L>& r, const Fe<2 * L>& a, const Fe<2 * L>& b) { sub_w<L>(f.b, r.l, a.l, b.l); sub_w<L>(f.b, r.l + L, a.l + L, b.l + L); }
// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u          (f2m_mul, build_f2m.js:152-194)
template <int L> static inline void mul(const Field2<L>& f, Fe<2 * L>& r, const Fe<2 * L>& a, const Fe<2 * L>& b) {
  uint64_t v0[L], v1[L], sa[L], sb[L], m[L];
  mul_w<L>(f.b, v0, a.l, b.l); mul_w<L>(f.b, v1, a.l + L, b.l + L);
  add_w<L>(f.b, sa, a.l, a.l + L); add_w<L>(f.b, sb, b.l, b.l + L); mul_w<L>(f.b, m, sa, sb);
  sub_w<L>(f.b, m, m, v0); sub_w<L>(f.b, r.l + L, m, v1); sub_w<L>(f.b, r.l, v0, v1);
}
// (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u                                             (f2m_square, build_f2m.js:290-330)
template <int L> static inline void sqr(const Field2<L>& f, Fe<2 * L>& r, const Fe<2 * L>& a) {
  uint64_t s[L], d[L], p[L];
  add_w<L>(f.b, s, a.l, a.l + L); sub_w<L>(f.b, d, a.l, a.l + L); mul_w<L>(f.b, p, a.l, a.l + L);
  mul_w<L>(f.b, r.l, s, d); add_w<L>(f.b, r.l + L, p, p);
}
template <class F> static inline void dbl(const F& f, Fe<F::W>& r, const Fe<F::W>& a) { add(f, r, a, a); }

template <int W> struct XYZZ { Fe<W> x, y, zz, zzz; };       // same layout as the device's XYZZ<C> (4 field elements)

template <int W> static inline bool is_inf(const XYZZ<W>& p) { return is_zero<W>(p.zz); }
template <class F> static inline void set_inf(const F& f, XYZZ<F::W>& p) {
  memset(&p, 0, sizeof p); for (int i = 0; i < F::W; i++) p.y.l[i] = f.one[i];
}
// dbl-2008-s-1 (a = 0)
template <class F> static inline void pdbl(const F& f, XYZZ<F::W>& r, const XYZZ<F::W>& p) {
  constexpr int W = F::W;
  if (is_inf<W>(p)) { r = p; return; }
  Fe<W> U, V, Wd, S, M, t, X3, Y3;
  dbl(f, U, p.y); sqr(f, V, U); mul(f, Wd, U, V); mul(f, S, p.x, V);
  sqr(f, t, p.x); dbl(f, M, t); add(f, M, M, t);
  sqr(f, X3, M); sub(f, X3, X3, S); sub(f, X3, X3, S);
  sub(f, t, S, X3); mul(f, t, M, t); mul(f, U, Wd, p.y); sub(f, Y3, t, U);
  Fe<W> zz, zzz; mul(f, zz, V, p.zz); mul(f, zzz, Wd, p.zzz);
  r.x = X3; r.y = Y3; r.zz = zz; r.zzz = zzz;
}
// add-2008-s, complete
template <class F> static inline void padd(const F& f, XYZZ<F::W>& acc, const XYZZ<F::W>& q) {
  constexpr int W = F::W;
  if (is_inf<W>(q)) return;
  if (is_inf<W>(acc)) { acc = q; return; }
  Fe<W> U1, U2, S1, S2, P, R, PP, PPP, Q, t;
  mul(f, U1, acc.x, q.zz); mul(f, U2, q.x, acc.zz);
  mul(f, S1, acc.y, q.zzz); mul(f, S2, q.y, acc.zzz);
  sub(f, P, U2, U1); sub(f, R, S2, S1);
  if (is_zero<W>(P)) {
    if (is_zero<W>(R)) { XYZZ<W> d; pdbl(f, d, q); acc = d; } else set_inf(f, acc);
    return;
  }
  sqr(f, PP, P); mul(f, PPP, P, PP); mul(f, Q, U1, PP);
  sqr(f, t, R); sub(f, t, t, PPP); sub(f, t, t, Q); sub(f, acc.x, t, Q);
  sub(f, t, Q, acc.x); mul(f, t, R, t); mul(f, Q, S1, PPP); sub(f, acc.y, t, Q);
  mul(f, t, acc.zz, q.zz); mul(f, acc.zz, t, PP);
  mul(f, t, acc.zzz, q.zzz); mul(f, acc.zzz, t, PPP);
}

// Incremental window combination.  folded: slots of (logB + 1) XYZZ points as written by k_gather_folded.
// result = sum_w 2^(off_w) * ( T_w[0] + sum_j 2^j T_w[2^j] ), off_w = bit offset of window w.  Every term is a point times a
// power of two, so ONE Horner pass over the bit positions does both the per-window sums and the window combination:
// acc = 2*acc + (terms with exponent e), e from the top down -- about nbits doublings and W*(logB+1) additions in total.
// Groups of slots are fed from the TOP window down (their exponent ranges do not interleave: a window's terms span
// off_w .. off_w + logB - 1 < off_(w+1)), so the host can consume a group while the GPU still works on the lower ones.
template <class F> struct Combiner {
  static constexpr int EW = F::W;
  F f; uint32_t W, Wd, c0, rem, logB, per; int cur;      // cur = lowest exponent already folded into acc (acc is scaled by 2^cur)
  XYZZ<EW> acc;
  uint32_t off(uint32_t w) const { return w * c0 + (w < rem ? w : rem); }
  void begin(const F& f_, uint32_t W_, uint32_t Wd_, uint32_t c0_, uint32_t rem_, uint32_t logB_) {
    f = f_; W = W_; Wd = Wd_; c0 = c0_; rem = rem_; logB = logB_; per = logB + 1;
    cur = (int)(off(Wd - 1) + logB + 1); set_inf(f, acc);
  }
  // slots [w0, w1) of the bucket array (w1 may be W, i.e. include the extra slot); `folded` points at slot 0
  void feed(const XYZZ<EW>* folded, uint32_t w0, uint32_t w1) {
    const uint32_t lastw = (w1 > Wd ? Wd : w1);                 // digit windows in this group: [w0, lastw)
    if (lastw <= w0) return;                                     // (a group holding only the extra slot cannot occur: it is cut with the last window)
    const int lo = (int)off(w0), hi = cur - 1;
    const uint32_t span = (uint32_t)(hi - lo + 1), nterms_max = (w1 - w0) * per + 2;
    const XYZZ<EW>** term = (const XYZZ<EW>**)__builtin_alloca(sizeof(void*) * nterms_max);
    int32_t* next = (int32_t*)__builtin_alloca(sizeof(int32_t) * nterms_max);
    int32_t* head = (int32_t*)__builtin_alloca(sizeof(int32_t) * span);
    for (uint32_t e = 0; e < span; e++) head[e] = -1;
    uint32_t nt = 0;
    auto put = [&](uint32_t e, const XYZZ<EW>* p) { if (is_inf<EW>(*p)) return; term[nt] = p; next[nt] = head[e - lo]; head[e - lo] = (int32_t)nt; nt++; };
    for (uint32_t w = w0; w < lastw; w++) {
      const XYZZ<EW>* T = folded + (size_t)w * per;
      put(off(w), &T[0]);
      for (uint32_t j = 0; j < logB; j++) put(off(w) + j, &T[1 + j]);
    }
    if (w1 > Wd) {   // extra slot: buckets B+1..2B of the last window: (2^logB + 1) E[0] + sum_j 2^j E[2^j]
      const XYZZ<EW>* E = folded + (size_t)Wd * per; const uint32_t o = off(Wd - 1);
      put(o, &E[0]); put(o + logB, &E[0]);
      for (uint32_t j = 0; j < logB; j++) put(o + j, &E[1 + j]);
    }
    for (int e = hi; e >= lo; e--) {
      if (!is_inf<EW>(acc)) { XYZZ<EW> d; pdbl(f, d, acc); acc = d; }
      for (int32_t t = head[e - lo]; t >= 0; t = next[t]) padd(f, acc, *term[t]);
    }
    cur = lo;
  }
  // out_jac: 3*EW words, Jacobian Montgomery
  void finish(uint64_t* out_jac) {
    for (; cur > 0; cur--) if (!is_inf<EW>(acc)) { XYZZ<EW> d; pdbl(f, d, acc); acc = d; }
    // XYZZ -> Jacobian without inversion: Z = ZZ*ZZZ, X = x*ZZ*ZZZ^2, Y = y*ZZ^3*ZZZ^2; infinity -> (0, R mod q, 0)
    Fe<EW> X, Y, Z;
    if (is_inf<EW>(acc)) { memset(&X, 0, sizeof X); memset(&Z, 0, sizeof Z); for (int i = 0; i < EW; i++) Y.l[i] = f.one[i]; }
    else {
      Fe<EW> t, u;
      mul(f, Z, acc.zz, acc.zzz); mul(f, t, Z, acc.zzz); mul(f, X, acc.x, t);
      sqr(f, u, acc.zz); mul(f, t, t, u); mul(f, Y, acc.y, t);
    }
    memcpy(out_jac, X.l, 8 * EW); memcpy(out_jac + EW, Y.l, 8 * EW); memcpy(out_jac + 2 * EW, Z.l, 8 * EW);
  }
};

// Window-table form (one bucket array cut into S sub-slots of 2^logBs buckets, each folded on its own):
//   sum_b (b+1) T[b] = sum_s V_s,   V_s = F_s[0] + sum_j 2^j F_s[2^j] + s * 2^logBs * F_s[0]      (F_s[0] = plain sum of sub-slot s)
// folded: (logBs + 1) XYZZ points per sub-slot as written by k_gather_folded.  Every V_s is one short Horner pass over its
// exponents (~logBs + log2 S doublings), independent of the other sub-slots, so groups are reduced as they arrive.
template <class F> struct SubslotCombiner {
  static constexpr int EW = F::W;
  F f; uint32_t S, logBs, per, sbits; XYZZ<EW> total;
  void begin(const F& f_, uint32_t S_, uint32_t logBs_) {
    f = f_; S = S_; logBs = logBs_; per = logBs + 1; sbits = 0; while ((1u << sbits) < S) sbits++;
    set_inf(f, total);
  }
  void feed(const XYZZ<EW>* folded, uint32_t s0, uint32_t s1) {
    for (uint32_t s = s0; s < s1; s++) {
      const XYZZ<EW>* Fs = folded + (size_t)s * per;
      const bool have0 = !is_inf<EW>(Fs[0]);
      XYZZ<EW> acc; set_inf(f, acc);
      for (int e = (int)(logBs + sbits) - 1; e >= 0; e--) {
        if (!is_inf<EW>(acc)) { XYZZ<EW> d; pdbl(f, d, acc); acc = d; }
        if ((uint32_t)e >= logBs) { if (have0 && ((s >> ((uint32_t)e - logBs)) & 1)) padd(f, acc, Fs[0]); }
        else if (!is_inf<EW>(Fs[1 + e])) padd(f, acc, Fs[1 + e]);
      }
      if (have0) padd(f, acc, Fs[0]);
      if (!is_inf<EW>(acc)) padd(f, total, acc);
    }
  }
  void finish(uint64_t* out_jac) { Combiner<F> cb; cb.f = f; cb.acc = total; cb.cur = 0; cb.finish(out_jac); }
};

}  // namespace b200host
