// accumulate.cuh -- bucket accumulation by batch-affine additions with Montgomery batch inversion.
//
// GPU redesign of the reference's opt-path bucket phase:
//   _constructAdditionChains / _evaluateAdditionChains / _addAffinePointsOneRound / _reduceBuckets
//   wasmcurves/src/build_multiexp_opt.js:651-858, :1016-1245, :1336-1585 and f1m_batchInverse,
//   wasmcurves/src/build_batchinverse.js:4-140.
// The reference groups each bucket's points by the binary expansion of its count and walks the
// levels serially.  Here every bucket is summed by a balanced pairwise tree: in round r every bucket
// segment of n_r points becomes ceil(n_r/2) points (adjacent points are added, an odd last point is
// carried over), for ALL buckets of ALL windows in one launch.  All additions of a round share ONE
// field inversion: a grid-wide product tree (K-ary, run-time K) of the denominators is built level by
// level, the single root is inverted by one thread, and the inverses are propagated back down.
//
//   per addition: forward 1M (running product) ; backward 2M (denominator inverse) + 2M + 1S (slope, x3, y3)
//   = 6 field multiplications = 6*(2N^2+N) limb products, the reference's own count
//   (build_multiexp_opt.js:1207-1233 + build_batchinverse.js:61-66,108-119).
//
// Segment bookkeeping: off[r][b] (exclusive scans of n_r[b] = ceil(n_0[b] / 2^r)) for every round are
// computed up front from the sort histogram; bid[r][j] gives the bucket of output slot j of round r-1
// and is propagated from round to round by the threads that own even local slots.
#pragma once
#include "msm_kernels.cuh"

namespace b200 {

constexpr int BA_THREADS = 128;     // threads per block in the tree kernels; a block owns a tile of K * BA_THREADS consecutive slots,
                                    // K = additions per thread per inversion chain (level 0) or product-tree arity (levels >= 1): run-time parameters
constexpr uint32_t BA_ROOT_MAX = 1024;   // = BA_ROOT_THREADS * ROOT_PER   // the product tree is reduced until at most this many values remain

// All rounds' segment offsets in one three-phase scan: off[r][b] = exclusive scan over b of ceil(n_0[b] / 2^r), r = 1..R.
// offs: R arrays of (n + 1) words; tile_sums: R arrays of (ntiles + 1) words.
__global__ void __launch_bounds__(SCAN_THREADS) k_mscan_tiles(const uint32_t* __restrict__ cnt0, uint32_t n, uint32_t R,
                                                              uint32_t* __restrict__ offs, uint32_t* __restrict__ tile_sums, uint32_t ntiles) {
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v0[SCAN_ITEMS];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) v0[k] = (base + k < n) ? cnt0[base + k] : 0;
  for (uint32_t r = 1; r <= R; r++) {
    uint32_t sum = 0, add = (1u << r) - 1u;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) sum += (v0[k] + add) >> r;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(sum, &total);
    uint32_t* out = offs + (size_t)(r - 1) * (n + 1);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += (v0[k] + add) >> r; }
    if (threadIdx.x == 0) tile_sums[(size_t)(r - 1) * (ntiles + 1) + blockIdx.x] = total;
  }
}
__global__ void k_mscan_sums(uint32_t* __restrict__ tile_sums, uint32_t ntiles) {      // one block per round
  uint32_t* ts = tile_sums + (size_t)blockIdx.x * (ntiles + 1);
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = (i < ntiles) ? ts[i] : 0, total;
    uint32_t ex = block_exclusive_scan(v, &total);
    uint32_t c = carry_s;
    if (i < ntiles) ts[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) ts[ntiles] = carry_s;
}
__global__ void k_mscan_apply(uint32_t* __restrict__ offs, uint32_t n, const uint32_t* __restrict__ tile_sums, uint32_t ntiles) {   // grid.y = round
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t* out = offs + (size_t)blockIdx.y * (n + 1);
  const uint32_t* ts = tile_sums + (size_t)blockIdx.y * (ntiles + 1);
  if (i < n) out[i] += ts[i / SCAN_TILE];
  if (i == 0) out[n] = ts[ntiles];
}

// bid1[j] = bucket owning output slot j of round 0.  One warp per 32 buckets: each lane fetches the slot range of its
// bucket, then the warp writes the 32 ranges one after the other with coalesced stores (a range is ~16 slots for
// uniform scalars; a heavy bucket's range is simply more iterations of the whole warp -- no per-slot binary search).
__global__ void __launch_bounds__(256) k_fill_bid(const uint32_t* __restrict__ off1, uint32_t nb, uint32_t* __restrict__ bid1) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t b = warp * 32 + lane;
  uint32_t lo = 0, hi = 0;
  if (b < nb) { lo = off1[b]; hi = off1[b + 1]; }
#pragma unroll 1
  for (int k = 0; k < 32; k++) {
    const uint32_t l = __shfl_sync(0xffffffffu, lo, k), h = __shfl_sync(0xffffffffu, hi, k);
    for (uint32_t j = l + lane; j < h; j += 32) bid1[j] = warp * 32 + k;
  }
}

// ---- one tree round, level 0 --------------------------------------------------------------------------
struct TreeRound {
  const uint32_t* off_in;     // off[r]    (nb + 1)
  const uint32_t* off_out;    // off[r+1]  (nb + 1)
  const uint32_t* off_next;   // off[r+2]  or nullptr
  const uint32_t* bid;        // bid[r+1]  (one per output slot)
  uint32_t* bid_next;         // bid[r+2]  or nullptr
  uint32_t nb;
};

// slot -> (input position of the first operand, has second operand)
B200_DI bool tree_slot(const TreeRound& tr, uint32_t j, uint32_t& in0, bool& has2, bool write_next) {
  uint32_t m = tr.off_out[tr.nb];
  if (j >= m) return false;
  uint32_t b = tr.bid[j];
  uint32_t local = j - tr.off_out[b];
  in0 = tr.off_in[b] + 2 * local;
  has2 = in0 + 1 < tr.off_in[b + 1];
  if (write_next && tr.bid_next && !(local & 1)) tr.bid_next[tr.off_next[b] + (local >> 1)] = b;
  return true;
}

// Per-slot operand table of a round, written once by k_tree_meta and read (coalesced) by the forward and the backward pass:
// meta[j] = (a, b): operand references of output slot j -- round 0: the sorted entries (point index | sign << 31), later rounds:
// positions in the previous round's output; b = NONE: the slot only carries its single input over; a = NONE: padding slot.
// This takes the dependent lookups (bid -> offsets -> sorted entry) out of the arithmetic kernels, whose gathers then depend
// on ONE coalesced load that is issued an iteration ahead; the table covers whole tiles, so those kernels need no bounds checks.
constexpr uint32_t META_NONE = 0xffffffffu;
template <bool FIRST>
__global__ void __launch_bounds__(256) k_tree_meta(TreeRound tr, const uint32_t* __restrict__ sorted, uint2* __restrict__ meta, uint32_t nslots) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nslots) return;
  uint32_t in0; bool has2;
  uint2 m = make_uint2(META_NONE, META_NONE);
  if (tree_slot(tr, j, in0, has2, true)) {
    if (FIRST) { m.x = __ldg(sorted + in0); if (has2) m.y = __ldg(sorted + in0 + 1); }
    else { m.x = in0; if (has2) m.y = in0 + 1; }
  }
  meta[j] = m;
}

// Round 0 reads the caller's bases (x || y records, 2*n8 bytes apart).  The rounds' own outputs are stored as TWO arrays, all x
// coordinates then (yoff bytes further) all y coordinates, because the forward pass needs only x: its sequential reads halve.
template <class C, bool FIRST>
B200_DI void meta_load_point(Affine<C>& p, const void* __restrict__ src, uint64_t yoff, uint32_t ref) {
  if (FIRST) { affine_load<C>(p, src, ref & 0x7fffffffu); if (ref >> 31) fe_neg<C>(p.y, p.y); }
  else { const char* b = reinterpret_cast<const char*>(src) + (uint64_t)ref * (4 * C::N); fe_load_cg<C>(p.x, b); fe_load_cg<C>(p.y, b + yoff); }
}
template <class C, bool FIRST>
B200_DI void meta_load_x(Fe<C::N>& x, const void* __restrict__ src, uint32_t ref) {
  if (FIRST) fe_load<C>(x, reinterpret_cast<const char*>(src) + (uint64_t)(ref & 0x7fffffffu) * (8 * C::N));
  else fe_load_cg<C>(x, reinterpret_cast<const char*>(src) + (uint64_t)ref * (4 * C::N));
}
template <class C>
B200_DI void soa_store_point(void* __restrict__ dst, uint64_t yoff, uint32_t j, const Affine<C>& p) {
  char* b = reinterpret_cast<char*>(dst) + (uint64_t)j * (4 * C::N); fe_store<C>(b, p.x); fe_store<C>(b + yoff, p.y);
}

// forward: denominators, per-slot prefix products, per-thread products.
// The kernel is gather-bound (two scattered 96-byte points per slot in round 0), so the x coordinates of slot i+1 are requested
// before the multiplication of slot i and the operand references of slot i+2 before that (only the x coordinates are needed;
// the y coordinates are fetched in the rare equal-x / zero-x cases).
// One tile (K * BA_THREADS consecutive slots, thread t owns slots t, t + 128, ...): returns the thread's running product in p.
template <class C, bool FIRST>
B200_DI void tree_fwd_tile(Fe<C::N>& p, const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff, void* __restrict__ prefix, int K, uint32_t tb) {
  const uint32_t tile = tb * (K * BA_THREADS) + threadIdx.x;
  fe_set_one<C>(p);
  Fe<C::N> x1, x2, nx1, nx2;
  uint2 mc = meta[tile], mn = K > 1 ? meta[tile + BA_THREADS] : make_uint2(META_NONE, META_NONE);
  if (mc.y != META_NONE) { meta_load_x<C, FIRST>(x1, src, mc.x); meta_load_x<C, FIRST>(x2, src, mc.y); }
#pragma unroll 1
  for (int i = 0; i < K; i++) {
    uint2 mn2 = make_uint2(META_NONE, META_NONE);
    if (i + 2 < K) mn2 = meta[tile + (i + 2) * BA_THREADS];
    if (i + 1 < K && mn.y != META_NONE) { meta_load_x<C, FIRST>(nx1, src, mn.x); meta_load_x<C, FIRST>(nx2, src, mn.y); }
    if (mc.y != META_NONE) {
      Fe<C::N> d; int kind = 0;
      fe_sub<C>(d, x2, x1);
      if (fe_is_zero<C>(d) || fe_is_zero<C>(x1) || fe_is_zero<C>(x2)) {        // rare: decide with the full points
        Affine<C> p1, p2;
        meta_load_point<C, FIRST>(p1, src, yoff, mc.x);
        meta_load_point<C, FIRST>(p2, src, yoff, mc.y);
        kind = affine_add_denominator<C>(d, p1, p2);
      }
      if (kind <= 1) {
        fe_store<C>(reinterpret_cast<char*>(prefix) + (uint64_t)(tile + i * BA_THREADS) * 4 * C::N, p);
        fe_mul<C>(p, p, d);
      }
    }
    mc = mn; mn = mn2; x1 = nx1; x2 = nx2;
  }
}
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 3 : 6) k_tree_fwd(const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff,
                                                            void* __restrict__ prefix, void* __restrict__ prod, int K, uint32_t ntiles) {
 // persistent form: gridDim.x may be smaller than ntiles (leaves SM room for the other lane's latency-bound kernels)
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> p;
  tree_fwd_tile<C, FIRST>(p, meta, src, yoff, prefix, K, tb);
  fe_store<C>(reinterpret_cast<char*>(prod) + (uint64_t)(tb * BA_THREADS + threadIdx.x) * 4 * C::N, p);
 }
}

// backward: consume the inverse q of the thread's product, finish every addition, write the round's output points
template <class C, bool FIRST>
B200_DI void tree_bwd_tile(Fe<C::N>& q, const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff, const void* __restrict__ prefix,
                           void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t tb) {
  const uint32_t tile = tb * (K * BA_THREADS) + threadIdx.x;
  uint2 mn = meta[tile + (K - 1) * BA_THREADS];
#pragma unroll 1
  for (int i = K - 1; i >= 0; i--) {
    const uint2 m = mn;
    if (i > 0) mn = meta[tile + (i - 1) * BA_THREADS];
    if (m.x == META_NONE) continue;
    const uint32_t j = tile + i * BA_THREADS;
    Affine<C> p1, p2, r;
    meta_load_point<C, FIRST>(p1, src, yoff, m.x);
    if (m.y == META_NONE) { soa_store_point<C>(pout, yoff_out, j, p1); continue; }
    meta_load_point<C, FIRST>(p2, src, yoff, m.y);
    Fe<C::N> d, dinv;
    int kind = affine_add_denominator<C>(d, p1, p2);
    if (kind <= 1) {
      Fe<C::N> pre; fe_load_cg<C>(pre, reinterpret_cast<const char*>(prefix) + (uint64_t)j * 4 * C::N);
      // the two multiplications of the inverse-sharing step are independent: their rows are interleaved (fe_mul2), which doubles the
      // work between dependent carry-chain instructions (measured: k_tree_bwd 3.46 -> 3.39 ms at 2^20, profiles/README.md r2)
      if constexpr (C::EXT == 1) { Fe<C::N> qn; fe_mul2<C>(dinv, q, pre, qn, q, d); q = qn; }
      else { fe_mul<C>(dinv, q, pre); fe_mul<C>(q, q, d); }
    }
    affine_add_finish<C>(r, p1, p2, dinv, kind);
    soa_store_point<C>(pout, yoff_out, j, r);
  }
}
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 2 : 4) k_tree_bwd(const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff,
                                                         const void* __restrict__ prefix, const void* __restrict__ inv,
                                                         void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles) {
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> q;
  fe_load_cg<C>(q, reinterpret_cast<const char*>(inv) + (uint64_t)(tb * BA_THREADS + threadIdx.x) * 4 * C::N);
  tree_bwd_tile<C, FIRST>(q, meta, src, yoff, prefix, pout, yoff_out, K, tb);
 }
}

#if defined(B200_EXPERIMENTS)      // round-2 measured alternatives (profiles/README.md r2): a register-lean backward pass and a single-launch tree round -- both slower, not shipped
// ---- backward pass, register-lean form ------------------------------------------------------------------------------------------
// The multiplier runs at the IMAD.WIDE peak from 8 warps per SM upwards, but a batch-affine addition is more than its five multiplications:
// gathers, carry chains of additions, stores.  While a warp is in those parts another must feed the pipe, and at 128 registers only four warps
// per scheduler are resident.  This form keeps NO operand alive across a multiplication: x1, x2 are loaded for the denominator and dropped,
// y1, y2 for the slope's numerator and dropped, and x1, x2, y1 are loaded AGAIN (L1 hits: the lines were touched moments ago) for x3 and y3.
// Live across the multiplications: q and the multiplication's own operands.  The rare cases (an operand at infinity, P + P, P + (-P)) go
// through the general code out of line.
template <class C, bool FIRST>
__device__ __noinline__ void bwd_slot_general(Fe<C::N>& q, uint2 m, const void* __restrict__ src, uint64_t yoff, const void* __restrict__ prefix,
                                              void* __restrict__ pout, uint64_t yoff_out, uint32_t j) {
  Affine<C> p1, p2, r;
  meta_load_point<C, FIRST>(p1, src, yoff, m.x);
  meta_load_point<C, FIRST>(p2, src, yoff, m.y);
  Fe<C::N> d, dinv;
  int kind = affine_add_denominator<C>(d, p1, p2);
  if (kind <= 1) {
    Fe<C::N> pre; fe_load_cg<C>(pre, reinterpret_cast<const char*>(prefix) + (uint64_t)j * 4 * C::N);
    fe_mul<C>(dinv, q, pre); fe_mul<C>(q, q, d);
  }
  affine_add_finish<C>(r, p1, p2, dinv, kind);
  soa_store_point<C>(pout, yoff_out, j, r);
}
template <class C, bool FIRST>
B200_DI void meta_load_y(Fe<C::N>& y, const void* __restrict__ src, uint64_t yoff, uint32_t ref) {
  if (FIRST) { fe_load<C>(y, reinterpret_cast<const char*>(src) + (uint64_t)(ref & 0x7fffffffu) * (8 * C::N) + 4 * C::N); if (ref >> 31) fe_neg<C>(y, y); }
  else fe_load_cg<C>(y, reinterpret_cast<const char*>(src) + (uint64_t)ref * (4 * C::N) + yoff);
}
B200_DI void reg_fence() { asm volatile("" ::: "memory"); }       // keeps the compiler from hoisting the later loads above the multiplications

template <class C, bool FIRST>
B200_DI void tree_bwd_tile_lean(Fe<C::N>& q, const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff, const void* __restrict__ prefix,
                                void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t tb) {
  const uint32_t tile = tb * (K * BA_THREADS) + threadIdx.x;
  uint2 mn = meta[tile + (K - 1) * BA_THREADS];
#pragma unroll 1
  for (int i = K - 1; i >= 0; i--) {
    const uint2 m = mn;
    if (i > 0) mn = meta[tile + (i - 1) * BA_THREADS];
    if (m.x == META_NONE) continue;
    const uint32_t j = tile + i * BA_THREADS;
    char* ox = reinterpret_cast<char*>(pout) + (uint64_t)j * (4 * C::N);
    if (m.y == META_NONE) { Affine<C> p1; meta_load_point<C, FIRST>(p1, src, yoff, m.x); soa_store_point<C>(pout, yoff_out, j, p1); continue; }
    Fe<C::N> d, t;
    { Fe<C::N> x1, x2; meta_load_x<C, FIRST>(x1, src, m.x); meta_load_x<C, FIRST>(x2, src, m.y);
      fe_sub<C>(d, x2, x1);
      if (fe_is_zero<C>(d) || fe_is_zero<C>(x1) || fe_is_zero<C>(x2)) { bwd_slot_general<C, FIRST>(q, m, src, yoff, prefix, pout, yoff_out, j); continue; } }
    { Fe<C::N> pre; fe_load_cg<C>(pre, reinterpret_cast<const char*>(prefix) + (uint64_t)j * 4 * C::N);
      fe_mul<C>(t, q, pre); }                        // t = 1 / d
    fe_mul<C>(q, q, d);
    reg_fence();
    { Fe<C::N> y1, y2; meta_load_y<C, FIRST>(y1, src, yoff, m.x); meta_load_y<C, FIRST>(y2, src, yoff, m.y);
      fe_sub<C>(d, y2, y1); }
    fe_mul<C>(t, d, t);                              // lambda
    fe_sqr<C>(d, t);
    reg_fence();
    { Fe<C::N> x1, x2; meta_load_x<C, FIRST>(x1, src, m.x); meta_load_x<C, FIRST>(x2, src, m.y);
      fe_sub<C>(d, d, x1); fe_sub<C>(d, d, x2);      // x3
      fe_store<C>(ox, d);
      fe_sub<C>(d, x1, d); }
    fe_mul<C>(t, t, d);
    reg_fence();
    { Fe<C::N> y1; meta_load_y<C, FIRST>(y1, src, yoff, m.x); fe_sub<C>(t, t, y1); }
    fe_store<C>(ox + yoff_out, t);
  }
}
#ifndef B200_BWD_LEAN_CTAS
#define B200_BWD_LEAN_CTAS 5
#endif
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, B200_BWD_LEAN_CTAS) k_tree_bwd_lean(const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff,
                                                         const void* __restrict__ prefix, const void* __restrict__ inv,
                                                         void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles) {
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> q;
  fe_load_cg<C>(q, reinterpret_cast<const char*>(inv) + (uint64_t)(tb * BA_THREADS + threadIdx.x) * 4 * C::N);
  tree_bwd_tile_lean<C, FIRST>(q, meta, src, yoff, prefix, pout, yoff_out, K, tb);
 }
}

// ---- one tree round in ONE persistent launch ------------------------------------------------------------------------------------
// k_tree_fwd, the product-tree levels, the root inversion and k_tree_bwd of a round as a single kernel of G co-resident CTAs.
// CTA c owns the tiles c, c + G, c + 2G, ... ("wave" w = the G tiles w*G .. w*G + G - 1).  Every wave is one batch inversion
// (f1m_batchInverse, build_batchinverse.js:4-140) of its own: a thread's running product goes up a 128-leaf tree in shared memory,
// the CTA's product goes to global memory, and one extra CTA (the "root CTA", which owns no tiles) multiplies the <= G CTA products of a
// wave together as soon as all of them are there (one chain per thread + the same shared-memory tree), inverts the single root (bingcd.h),
// walks back down and releases the wave.  (A first version let the LAST-arriving worker run the root: that worker then fell one root behind
// per wave and every wave waited for it -- 3.8 instead of 2.1 ms for round 0 at 2^20.)
// The waves are software-pipelined: a CTA runs the forward passes of waves w+1 and w+2 BEFORE it waits for the inverse of wave w, so the
// latency of a wave's root (tree + one inversion, ~70 us) is covered by forward work and by the other CTAs of
// the SM, which drift out of phase -- gather-bound forward tiles and multiplier-bound backward tiles then share an SM, instead of
// running as separate kernels one after the other.  Launched cooperatively (all G CTAs resident: the waits below cannot starve), with
// a clock-bounded spin as a backstop that raises sy.error instead of hanging the device.
struct RoundSync {
  uint32_t* arrive;      // [waves]      CTAs that have published their product of wave w
  uint32_t* ready;       // [waves]      1 once cta_inv[w][*] is complete
  void* cta_prod;        // [waves][G]   CTA products
  void* cta_inv;         // [waves][G]   their inverses (scratch for the prefix products of the root's chains before that)
  uint32_t* error;       // [1]          a spin wait timed out (results invalid)
};
constexpr long long ROUND_SPIN_LIMIT = 4000000000ll;      // ~2 s at 1.9 GHz

B200_DI uint32_t ld_volatile_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// 128-leaf product tree in shared memory, heap order (node 1 = root, leaves at [T, 2T)); all threads of the CTA call these
template <class C> B200_DI void cta_tree_up(uint32_t* __restrict__ tree, const Fe<C::N>& leaf) {
  constexpr int N = C::N; const uint32_t t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < N; k++) tree[(BA_THREADS + t) * N + k] = leaf.l[k];
  __syncthreads();
#pragma unroll 1
  for (uint32_t width = BA_THREADS / 2; width >= 1; width >>= 1) {
    if (t < width) {
      Fe<N> a, b, c; const uint32_t node = width + t;
#pragma unroll
      for (int k = 0; k < N; k++) { a.l[k] = tree[(2 * node) * N + k]; b.l[k] = tree[(2 * node + 1) * N + k]; }
      fe_mul<C>(c, a, b);
#pragma unroll
      for (int k = 0; k < N; k++) tree[node * N + k] = c.l[k];
    }
    __syncthreads();
  }
}
// node 1 holds the inverse of the root on entry; on exit q = inverse of this thread's leaf
template <class C> B200_DI void cta_tree_down(uint32_t* __restrict__ tree, Fe<C::N>& q) {
  constexpr int N = C::N; const uint32_t t = threadIdx.x;
#pragma unroll 1
  for (uint32_t width = 1; width < BA_THREADS; width <<= 1) {
    if (t < width) {
      Fe<N> a, b, ip, ia, ib; const uint32_t node = width + t;
#pragma unroll
      for (int k = 0; k < N; k++) { ip.l[k] = tree[node * N + k]; a.l[k] = tree[(2 * node) * N + k]; b.l[k] = tree[(2 * node + 1) * N + k]; }
      fe_mul<C>(ia, ip, b); fe_mul<C>(ib, ip, a);
#pragma unroll
      for (int k = 0; k < N; k++) { tree[(2 * node) * N + k] = ia.l[k]; tree[(2 * node + 1) * N + k] = ib.l[k]; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < N; k++) q.l[k] = tree[(BA_THREADS + t) * N + k];
}
// the wave's root, run by the whole CTA that arrived last: cnt CTA products -> their inverses
template <class C> __device__ __noinline__ void round_root(uint32_t* __restrict__ tree, const void* __restrict__ prods, void* __restrict__ invs, uint32_t cnt) {
  constexpr int N = C::N; const uint32_t t = threadIdx.x;
  Fe<N> p; fe_set_one<C>(p);
#pragma unroll 1
  for (uint32_t e = t; e < cnt; e += BA_THREADS) {       // chain of this thread; its prefix products wait in invs[]
    Fe<N> v; fe_load_l2<C>(v, reinterpret_cast<const char*>(prods) + (uint64_t)e * 4 * N);
    fe_store<C>(reinterpret_cast<char*>(invs) + (uint64_t)e * 4 * N, p);
    fe_mul<C>(p, p, v);
  }
  cta_tree_up<C>(tree, p);
  if (t == 0) {
    Fe<N> r, ri;
#pragma unroll
    for (int k = 0; k < N; k++) r.l[k] = tree[1 * N + k];
    fe_inv_fast<C>(ri, r);
#pragma unroll
    for (int k = 0; k < N; k++) tree[1 * N + k] = ri.l[k];
  }
  __syncthreads();
  Fe<N> q; cta_tree_down<C>(tree, q);
  if (cnt > 0) {
    const uint32_t last = t + ((cnt - 1 - t) / BA_THREADS) * BA_THREADS;      // largest e = t (mod 128) below cnt (t < cnt)
    if (t < cnt) {
#pragma unroll 1
      for (int64_t e = last; e >= (int64_t)t; e -= BA_THREADS) {
        Fe<N> v, pre, r;
        fe_load_l2<C>(v, reinterpret_cast<const char*>(prods) + (uint64_t)e * 4 * N);
        fe_load_l2<C>(pre, reinterpret_cast<const char*>(invs) + (uint64_t)e * 4 * N);
        fe_mul<C>(r, q, pre); fe_mul<C>(q, q, v);
        fe_store<C>(reinterpret_cast<char*>(invs) + (uint64_t)e * 4 * N, r);
      }
    }
  }
}

constexpr uint32_t ROUND_DEPTH = 2;       // waves whose forward pass runs ahead of the backward pass (tree buffers: ROUND_DEPTH + 1)

// gridDim.x = G + 1: CTAs 0..G-1 are the workers, the last CTA only runs the waves' roots (so that no worker falls behind by a root per wave)
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, 4) k_tree_round(const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff, void* __restrict__ prefix,
                                                              void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles, RoundSync sy) {
  constexpr int N = C::N;
  __shared__ uint32_t trees[ROUND_DEPTH + 1][2 * BA_THREADS * N];       // the trees of the waves in flight (the root CTA uses the first)
  __shared__ uint32_t s_stop;
  const uint32_t G = gridDim.x - 1, c = blockIdx.x, t = threadIdx.x;
  const uint32_t nw = (ntiles + G - 1) / G;
  auto spin_until = [&](const uint32_t* flag, uint32_t want) {          // thread 0 only; false = gave up (error raised by this or another CTA)
    const long long t0 = clock64();
    while (ld_volatile_u32(flag) < want) {
      __nanosleep(64);
      if (ld_volatile_u32(sy.error) != 0u) return false;
      if (clock64() - t0 > ROUND_SPIN_LIMIT) { atomicExch(sy.error, 1u); return false; }
    }
    return true;
  };
  if (c == G) {
    // ---- root CTA: wave after wave, wait until all its CTAs have published, invert, release
#pragma unroll 1
    for (uint32_t w = 0; w < nw; w++) {
      const uint32_t cnt = min(G, ntiles - w * G);
      if (t == 0) { s_stop = spin_until(sy.arrive + w, cnt) ? 0u : 1u; __threadfence(); }
      __syncthreads();
      if (s_stop) return;
      round_root<C>(trees[0], reinterpret_cast<const char*>(sy.cta_prod) + (uint64_t)w * G * 4 * N, reinterpret_cast<char*>(sy.cta_inv) + (uint64_t)w * G * 4 * N, cnt);
      __threadfence();
      __syncthreads();
      if (t == 0) atomicExch(sy.ready + w, 1u);
    }
    return;
  }
  if (c >= ntiles) return;
  const uint32_t mine = (ntiles - c + G - 1) / G;          // waves this CTA takes part in (>= 1)
#pragma unroll 1
  for (uint32_t it = 0; it < mine + ROUND_DEPTH; it++) {   // iteration `it`: forward pass of wave `it`, then wait + backward pass of wave `it - ROUND_DEPTH`
    if (it < mine) {
      Fe<N> p;
      tree_fwd_tile<C, FIRST>(p, meta, src, yoff, prefix, K, it * G + c);
      uint32_t* tr = trees[it % (ROUND_DEPTH + 1)];
      cta_tree_up<C>(tr, p);
      if (t == 0) {
        Fe<N> r;
#pragma unroll
        for (int k = 0; k < N; k++) r.l[k] = tr[1 * N + k];
        fe_store<C>(reinterpret_cast<char*>(sy.cta_prod) + ((uint64_t)it * G + c) * 4 * N, r);
        __threadfence();
        atomicAdd(sy.arrive + it, 1u);
      }
    }
    if (it < ROUND_DEPTH) continue;
    const uint32_t w = it - ROUND_DEPTH;
    uint32_t* tr = trees[w % (ROUND_DEPTH + 1)];
    if (t == 0) {
      spin_until(sy.ready + w, 1u);
      __threadfence();
      Fe<N> r; fe_load_l2<C>(r, reinterpret_cast<const char*>(sy.cta_inv) + ((uint64_t)w * G + c) * 4 * N);
#pragma unroll
      for (int k = 0; k < N; k++) tr[1 * N + k] = r.l[k];
    }
    __syncthreads();
    Fe<N> q; cta_tree_down<C>(tr, q);
    tree_bwd_tile<C, FIRST>(q, meta, src, yoff, prefix, pout, yoff_out, K, w * G + c);
    __syncthreads();                                       // the tree buffer of wave w is free for wave w + ROUND_DEPTH + 1
  }
}

#endif  // B200_EXPERIMENTS

#if defined(B200_EXPERIMENTS)      // measured 0-8 % slower than k_tree_bwd (profiles/README.md): not in the shipped library
// backward pass with operand staging: the 2 points + prefix product of the NEXT slot are copied global -> shared with cp.async
// while the current slot's five multiplications run, so the arithmetic never waits on a gather (the plain kernel above shows
// 2.3 of its 4 warps per scheduler stalled on the scoreboard).  Each thread stages only its own operands (chunk-major layout:
// 16-byte chunk c of thread t at sm[c * BA_THREADS + t], conflict-free), so no CTA barrier is involved: cp.async.wait_group
// orders a thread's own copies, and the next copy is issued only after the registers read from the buffer have been consumed.
B200_DI void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
B200_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
B200_DI void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <class C, bool FIRST>
B200_DI void bwd_stage(uint4* sm, uint2 m, const void* __restrict__ src, uint64_t yoff, const void* __restrict__ prefix, uint32_t j) {
  constexpr int FC = C::N / 4, PC = 2 * FC;       // 16-byte chunks per field element / per point
  if (m.x != META_NONE) {
    const char* b1 = reinterpret_cast<const char*>(src) + (FIRST ? (uint64_t)(m.x & 0x7fffffffu) * (8 * C::N) : (uint64_t)m.x * (4 * C::N));
#pragma unroll
    for (int c = 0; c < PC; c++) cp_async16(sm + c * BA_THREADS + threadIdx.x, (FIRST || c < FC) ? b1 + 16 * c : b1 + yoff + 16 * (c - FC));
    if (m.y != META_NONE) {
      const char* b2 = reinterpret_cast<const char*>(src) + (FIRST ? (uint64_t)(m.y & 0x7fffffffu) * (8 * C::N) : (uint64_t)m.y * (4 * C::N));
      const uint4* gp = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(prefix) + (uint64_t)j * (4 * C::N));
#pragma unroll
      for (int c = 0; c < PC; c++) cp_async16(sm + (PC + c) * BA_THREADS + threadIdx.x, (FIRST || c < FC) ? b2 + 16 * c : b2 + yoff + 16 * (c - FC));
#pragma unroll
      for (int c = 0; c < FC; c++) cp_async16(sm + (2 * PC + c) * BA_THREADS + threadIdx.x, gp + c);
    }
  }
  cp_async_commit();
}
template <class C> B200_DI void sm_read_fe(Fe<C::N>& r, const uint4* sm, int chunk0) {
#pragma unroll
  for (int c = 0; c < C::N / 4; c++) { const uint4 v = sm[(chunk0 + c) * BA_THREADS + threadIdx.x]; r.l[4 * c] = v.x; r.l[4 * c + 1] = v.y; r.l[4 * c + 2] = v.z; r.l[4 * c + 3] = v.w; }
}

template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 2 : 4) k_tree_bwd_staged(const uint2* __restrict__ meta, const void* __restrict__ src, uint64_t yoff,
                                                                const void* __restrict__ prefix, const void* __restrict__ inv,
                                                                void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles) {
 extern __shared__ uint4 sm[];
 constexpr int FC = C::N / 4, PC = 2 * FC;
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  const uint32_t tile = tb * (K * BA_THREADS) + threadIdx.x;
  uint2 mn = meta[tile + (K - 1) * BA_THREADS];
  bwd_stage<C, FIRST>(sm, mn, src, yoff, prefix, tile + (K - 1) * BA_THREADS);
  Fe<C::N> q;
  fe_load_cg<C>(q, reinterpret_cast<const char*>(inv) + (uint64_t)(tb * BA_THREADS + threadIdx.x) * 4 * C::N);
#pragma unroll 1
  for (int i = K - 1; i >= 0; i--) {
    const uint2 m = mn;
    mn = make_uint2(META_NONE, META_NONE);
    if (i > 0) mn = meta[tile + (i - 1) * BA_THREADS];
    const uint32_t j = tile + i * BA_THREADS;
    cp_async_wait_all();
    if (m.x == META_NONE) { bwd_stage<C, FIRST>(sm, mn, src, yoff, prefix, j - BA_THREADS); continue; }
    Affine<C> p1, p2, r;
    sm_read_fe<C>(p1.x, sm, 0); sm_read_fe<C>(p1.y, sm, FC);
    if (FIRST && (m.x >> 31)) fe_neg<C>(p1.y, p1.y);
    if (m.y == META_NONE) {
      soa_store_point<C>(pout, yoff_out, j, p1);       // the store consumes the registers read from the buffer
      bwd_stage<C, FIRST>(sm, mn, src, yoff, prefix, j - BA_THREADS);
      continue;
    }
    sm_read_fe<C>(p2.x, sm, PC); sm_read_fe<C>(p2.y, sm, PC + FC);
    if (FIRST && (m.y >> 31)) fe_neg<C>(p2.y, p2.y);
    Fe<C::N> d, dinv, pre;
    sm_read_fe<C>(pre, sm, 2 * PC);
    int kind = affine_add_denominator<C>(d, p1, p2);
    if (kind <= 1) {
      fe_mul<C>(dinv, q, pre);
      fe_mul<C>(q, q, d);
    }
    bwd_stage<C, FIRST>(sm, mn, src, yoff, prefix, j - BA_THREADS);     // p1, p2, pre are in registers (consumed above): the buffer is free
    affine_add_finish<C>(r, p1, p2, dinv, kind);
    soa_store_point<C>(pout, yoff_out, j, r);
  }
  cp_async_wait_all();
 }
}

#endif  // B200_EXPERIMENTS

// ---- product tree, levels >= 1: plain arrays of field elements -----------------------------------------
// A level reduces n values by K per thread (serial running product, prefixes stored) and, when WARP is set, by a further
// factor 32 inside each warp: inclusive prefix and suffix products across the lanes by shuffles (5 + 5 steps that
// interleave), "others" = product of all other lanes' values is stored per thread, and the warp total goes up.
// Going back down, a thread's inverse is (inverse of the warp total) * others -- one multiplication.
// Small levels are latency-bound, so trading a few redundant multiplications for a 32x larger arity removes launches.
template <class C> B200_DI void fe_shfl_up(Fe<C::N>& r, const Fe<C::N>& a, int o) {
#pragma unroll
  for (int k = 0; k < C::N; k++) r.l[k] = __shfl_up_sync(0xffffffffu, a.l[k], o);
}
template <class C> B200_DI void fe_shfl_down(Fe<C::N>& r, const Fe<C::N>& a, int o) {
#pragma unroll
  for (int k = 0; k < C::N; k++) r.l[k] = __shfl_down_sync(0xffffffffu, a.l[k], o);
}

template <class C, bool WARP>
__global__ void __launch_bounds__(BA_THREADS) k_prod_fwd(const void* __restrict__ vals, uint32_t n, void* __restrict__ prefix, void* __restrict__ prod,
                                                         void* __restrict__ others, int K) {
  uint32_t tile = blockIdx.x * (K * BA_THREADS);
  Fe<C::N> p; fe_set_one<C>(p);
#pragma unroll 1
  for (int i = 0; i < K; i++) {
    uint32_t e = tile + i * BA_THREADS + threadIdx.x;
    if (e >= n) continue;
    Fe<C::N> v; fe_load_cg<C>(v, reinterpret_cast<const char*>(vals) + (uint64_t)e * 4 * C::N);
    fe_store<C>(reinterpret_cast<char*>(prefix) + (uint64_t)e * 4 * C::N, p);
    fe_mul<C>(p, p, v);
  }
  const uint32_t T = blockIdx.x * BA_THREADS + threadIdx.x;
  if (!WARP) { fe_store<C>(reinterpret_cast<char*>(prod) + (uint64_t)T * 4 * C::N, p); return; }
  const uint32_t lane = threadIdx.x & 31;
  Fe<C::N> pre = p, suf = p, y, z;
#pragma unroll 1
  for (int o = 1; o < 32; o <<= 1) {
    fe_shfl_up<C>(y, pre, o); fe_shfl_down<C>(z, suf, o);
    if (lane >= (uint32_t)o) fe_mul<C>(pre, pre, y);
    if (lane + o < 32) fe_mul<C>(suf, suf, z);
  }
  fe_shfl_up<C>(y, pre, 1); fe_shfl_down<C>(z, suf, 1);      // exclusive prefix / suffix
  Fe<C::N> oth;
  if (lane == 0) oth = z; else if (lane == 31) oth = y; else fe_mul<C>(oth, y, z);
  fe_store<C>(reinterpret_cast<char*>(others) + (uint64_t)T * 4 * C::N, oth);
  if (lane == 31) fe_store<C>(reinterpret_cast<char*>(prod) + (uint64_t)(T >> 5) * 4 * C::N, pre);
}
// in place: vals[e] <- 1 / vals[e], given the inverse of each thread's (or warp's) product in inv[]
template <class C, bool WARP>
__global__ void __launch_bounds__(BA_THREADS) k_prod_bwd(void* __restrict__ vals, uint32_t n, const void* __restrict__ prefix, const void* __restrict__ inv,
                                                         const void* __restrict__ others, int K) {
  uint32_t tile = blockIdx.x * (K * BA_THREADS);
  const uint32_t T = blockIdx.x * BA_THREADS + threadIdx.x;
  Fe<C::N> q;
  if (WARP) {
    Fe<C::N> wi, oth;
    fe_load_cg<C>(wi, reinterpret_cast<const char*>(inv) + (uint64_t)(T >> 5) * 4 * C::N);
    fe_load_cg<C>(oth, reinterpret_cast<const char*>(others) + (uint64_t)T * 4 * C::N);
    fe_mul<C>(q, wi, oth);
  } else {
    fe_load_cg<C>(q, reinterpret_cast<const char*>(inv) + (uint64_t)T * 4 * C::N);
  }
#pragma unroll 1
  for (int i = K - 1; i >= 0; i--) {
    uint32_t e = tile + i * BA_THREADS + threadIdx.x;
    if (e >= n) continue;
    Fe<C::N> v, pre, r;
    fe_load_cg<C>(v, reinterpret_cast<const char*>(vals) + (uint64_t)e * 4 * C::N);
    fe_load_cg<C>(pre, reinterpret_cast<const char*>(prefix) + (uint64_t)e * 4 * C::N);
    fe_mul<C>(r, q, pre);
    fe_mul<C>(q, q, v);
    fe_store<C>(reinterpret_cast<char*>(vals) + (uint64_t)e * 4 * C::N, r);
  }
}
// root: n <= BA_ROOT_MAX values inverted in place by ONE block of BA_ROOT_THREADS threads: each thread forms the running
// product of up to ROOT_PER values, a binary product tree over the threads lives in shared memory (log depth), thread 0
// inverts the root once (f1m_inverse at build_batchinverse.js:90), then the tree and the per-thread chains are walked back.
// (256 threads x 4 values; 128 x 8 for the 24-limb Fq2 of BLS12-381 G2, whose tree would not fit in 48 KB of static shared memory otherwise)
template <class C> struct RootCfg { static constexpr uint32_t THREADS = C::N > 16 ? 128 : 256, PER = BA_ROOT_MAX / THREADS; };
template <class C>
__global__ void __launch_bounds__(RootCfg<C>::THREADS) k_inv_root(void* __restrict__ vals, uint32_t n) {
  constexpr int N = C::N;
  constexpr uint32_t BA_ROOT_THREADS = RootCfg<C>::THREADS, ROOT_PER = RootCfg<C>::PER;
  __shared__ uint32_t tree[2 * BA_ROOT_THREADS * N];       // node k (1-based heap order): leaves at [BA_ROOT_THREADS, 2*BA_ROOT_THREADS)
  const uint32_t t = threadIdx.x;
  Fe<N> v[ROOT_PER], pre[ROOT_PER], p;
  fe_set_one<C>(p);
#pragma unroll
  for (uint32_t i = 0; i < ROOT_PER; i++) {
    uint32_t e = i * BA_ROOT_THREADS + t;
    if (e < n) fe_load_cg<C>(v[i], reinterpret_cast<const char*>(vals) + (uint64_t)e * 4 * N); else fe_set_one<C>(v[i]);
    pre[i] = p;
    if (e < n) fe_mul<C>(p, p, v[i]);
  }
#pragma unroll
  for (int k = 0; k < N; k++) tree[(BA_ROOT_THREADS + t) * N + k] = p.l[k];
  __syncthreads();
  for (uint32_t width = BA_ROOT_THREADS / 2; width >= 1; width >>= 1) {     // up-sweep: node = left * right
    if (t < width) {
      Fe<N> a, b, c; uint32_t node = width + t;
#pragma unroll
      for (int k = 0; k < N; k++) { a.l[k] = tree[(2 * node) * N + k]; b.l[k] = tree[(2 * node + 1) * N + k]; }
      fe_mul<C>(c, a, b);
#pragma unroll
      for (int k = 0; k < N; k++) tree[node * N + k] = c.l[k];
    }
    __syncthreads();
  }
  if (t == 0) {
    Fe<N> r, ri;
#pragma unroll
    for (int k = 0; k < N; k++) r.l[k] = tree[1 * N + k];
    fe_inv_fast<C>(ri, r);
#pragma unroll
    for (int k = 0; k < N; k++) tree[1 * N + k] = ri.l[k];
  }
  __syncthreads();
  for (uint32_t width = 1; width < BA_ROOT_THREADS; width <<= 1) {          // down-sweep: inv(left) = inv(node) * right, inv(right) = inv(node) * left
    if (t < width) {
      Fe<N> a, b, ip, ia, ib; uint32_t node = width + t;
#pragma unroll
      for (int k = 0; k < N; k++) { ip.l[k] = tree[node * N + k]; a.l[k] = tree[(2 * node) * N + k]; b.l[k] = tree[(2 * node + 1) * N + k]; }
      fe_mul<C>(ia, ip, b); fe_mul<C>(ib, ip, a);
#pragma unroll
      for (int k = 0; k < N; k++) { tree[(2 * node) * N + k] = ia.l[k]; tree[(2 * node + 1) * N + k] = ib.l[k]; }
    }
    __syncthreads();
  }
  Fe<N> q;
#pragma unroll
  for (int k = 0; k < N; k++) q.l[k] = tree[(BA_ROOT_THREADS + t) * N + k];
#pragma unroll
  for (int i = (int)ROOT_PER - 1; i >= 0; i--) {
    uint32_t e = (uint32_t)i * BA_ROOT_THREADS + t;
    if (e < n) {
      Fe<N> r; fe_mul<C>(r, q, pre[i]); fe_mul<C>(q, q, v[i]);
      fe_store<C>(reinterpret_cast<char*>(vals) + (uint64_t)e * 4 * N, r);
    }
  }
}

// ---- finish: one thread per bucket sums what is left of its segment and writes the bucket as XYZZ -------
template <class C, bool FIRST>
__global__ void __launch_bounds__(128) k_accum_finish(const void* __restrict__ bases, const uint32_t* __restrict__ sorted, const void* __restrict__ pin, uint64_t yoff,
                                                      const uint32_t* __restrict__ off, uint32_t nb, void* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t lo = off[b], hi = off[b + 1];
  XYZZ<C> acc; xyzz_set_inf<C>(acc);
  for (uint32_t k = lo; k < hi; k++) {
    Affine<C> p;
    if (FIRST) meta_load_point<C, true>(p, bases, 0, __ldg(sorted + k)); else meta_load_point<C, false>(p, pin, yoff, k);
    xyzz_madd<C>(acc, p);
  }
  xyzz_store<C>(buckets, b, acc);
}

}  // namespace b200
