// host_ec_harness.cpp -- CPU test harness for zprize-wasm-msm_b200/csrc/host_ec.h (the serial tail of the MSM).
// Built by tests/test_host_tail.py with g++ as a shared library; plain C entry points over byte buffers so that the
// Python side can feed folded-bucket arrays built with big-integer arithmetic and compare the Jacobian result.
//   ext = 1: prime field Fq (G1), ext = 2: Fq2 = Fq[u]/(u^2+1) (G2).  words = 64-bit words per Fq element (6 or 4).
#include <stdint.h>
#include <string.h>
#include "host_ec.h"

using namespace b200host;

template <int L> static Field<L> make_field(const uint64_t* q, const uint64_t* one) {
  Field<L> f; for (int i = 0; i < L; i++) { f.q[i] = q[i]; f.one[i] = one[i]; }
  uint64_t x = 1; for (int k = 0; k < 6; k++) x *= 2 - f.q[0] * x; f.np = 0 - x; return f;
}
template <int L> static Field2<L> make_field2(const uint64_t* q, const uint64_t* one) {
  Field2<L> f; f.b = make_field<L>(q, one); for (int i = 0; i < 2 * L; i++) f.one[i] = i < L ? one[i] : 0; return f;
}
template <class F> static void run_windows(const F& f, const void* folded, uint32_t W, uint32_t Wd, uint32_t c0, uint32_t rem, uint32_t logB, uint32_t split, void* out) {
  Combiner<F> cb; cb.begin(f, W, Wd, c0, rem, logB);
  const XYZZ<F::W>* p = reinterpret_cast<const XYZZ<F::W>*>(folded);
  if (split == 0 || split >= W) cb.feed(p, 0, W);
  else { cb.feed(p, split, W); cb.feed(p, 0, split); }          // groups arrive from the top windows down
  cb.finish(reinterpret_cast<uint64_t*>(out));
}
template <class F> static void run_subslots(const F& f, const void* folded, uint32_t S, uint32_t logBs, uint32_t split, void* out) {
  SubslotCombiner<F> cb; cb.begin(f, S, logBs);
  const XYZZ<F::W>* p = reinterpret_cast<const XYZZ<F::W>*>(folded);
  if (split == 0 || split >= S) cb.feed(p, 0, S); else { cb.feed(p, split, S); cb.feed(p, 0, split); }
  cb.finish(reinterpret_cast<uint64_t*>(out));
}

extern "C" {
// mode 0: Combiner (window slots), mode 1: SubslotCombiner.  a..e: (W, Wd, c0, rem, logB) or (S, logBs, -, -, -).  Returns 0, or -1 for bad sizes.
int host_tail(int ext, int words, const uint64_t* q, const uint64_t* one, int mode, const void* folded,
              uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t split, void* out) {
#define DISPATCH(FIELD) do { if (mode == 0) run_windows(FIELD, folded, a, b, c, d, e, split, out); else run_subslots(FIELD, folded, a, b, split, out); return 0; } while (0)
  if (ext == 1 && words == 6) DISPATCH(make_field<6>(q, one));
  if (ext == 1 && words == 4) DISPATCH(make_field<4>(q, one));
  if (ext == 2 && words == 6) DISPATCH(make_field2<6>(q, one));
  if (ext == 2 && words == 4) DISPATCH(make_field2<4>(q, one));
  return -1;
}
// r = a * b / r = a^2 in the field (Montgomery form), for the field-level check
int host_mul(int ext, int words, const uint64_t* q, const uint64_t* one, const void* x, const void* y, void* r, int square) {
#define MULCASE(L, F, W_) do { auto f = F; Fe<W_> a_, b_, r_; memcpy(&a_, x, sizeof a_); memcpy(&b_, y, sizeof b_); \
    if (square) sqr(f, r_, a_); else mul(f, r_, a_, b_); memcpy(r, &r_, sizeof r_); return 0; } while (0)
  if (ext == 1 && words == 6) MULCASE(6, make_field<6>(q, one), 6);
  if (ext == 1 && words == 4) MULCASE(4, make_field<4>(q, one), 4);
  if (ext == 2 && words == 6) MULCASE(6, make_field2<6>(q, one), 12);
  if (ext == 2 && words == 4) MULCASE(4, make_field2<4>(q, one), 8);
  return -1;
}
}
