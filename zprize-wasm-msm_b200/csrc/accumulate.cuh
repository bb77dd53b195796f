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

#ifndef B200_BWD_CTAS_N8
#define B200_BWD_CTAS_N8 5      // resident CTAs per SM k_tree_bwd is compiled for on the 8-limb field (BN254): 94 registers, no spill; measured against 4 (106 registers):
                                // k_tree_bwd 1.551 -> 1.510 ms at 2^20, the MSM 3.342 -> 3.294, 2^22 10.97 -> 10.84 (profiles/README.md r2late).  BLS12-381 needs 128 registers: 4.
#endif
#ifndef B200_FWD_CTAS_N12
#define B200_FWD_CTAS_N12 5     // k_tree_fwd on the 12-limb field (BLS12-381): 96 registers, no spill, against 80 registers + 64-96 bytes of spill at 6 CTAs per SM.  With the
                                // x-only copy of the bases the pass is issue-bound rather than HBM-bound: 0.997 -> 0.854 ms at 2^20, the MSM 6.25 -> 6.17, 2^22 20.04 -> 19.73 (r2late)
#endif
#ifndef B200_FWD_CTAS_N8
#define B200_FWD_CTAS_N8 6      // (7 measured equal: 72 registers either way)
#endif
constexpr int BA_THREADS = 128;     // threads per block in the tree kernels; a block owns a tile of K * BA_THREADS consecutive slots,
                                    // K = additions per thread per inversion chain (level 0) or product-tree arity (levels >= 1): run-time parameters
constexpr uint32_t BA_ROOT_MAX = 1024;   // = BA_ROOT_THREADS * ROOT_PER   // the product tree is reduced until at most this many values remain

// All rounds' segment offsets in one three-phase scan: off[r][b] = exclusive scan over b of ceil(n_0[b] / 2^r), r = 1..R.
// offs: R arrays of (n + 1) words; tile_sums: R arrays of (ntiles + 1) words.
B200_KERNEL void __launch_bounds__(SCAN_THREADS) k_mscan_tiles(const uint32_t* __restrict__ cnt0, uint32_t n, uint32_t R,
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
B200_KERNEL void k_mscan_sums(uint32_t* __restrict__ tile_sums, uint32_t ntiles) {      // one block per round
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
B200_KERNEL void k_mscan_apply(uint32_t* __restrict__ offs, uint32_t n, const uint32_t* __restrict__ tile_sums, uint32_t ntiles) {   // grid.y = round
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t* out = offs + (size_t)blockIdx.y * (n + 1);
  const uint32_t* ts = tile_sums + (size_t)blockIdx.y * (ntiles + 1);
  if (i < n) out[i] += ts[i / SCAN_TILE];
  if (i == 0) out[n] = ts[ntiles];
}

// bid1[j] = bucket owning output slot j of round 0.  One warp per 32 buckets: each lane fetches the slot range of its
// bucket, then the warp writes the 32 ranges one after the other with coalesced stores (a range is ~16 slots for
// uniform scalars; a heavy bucket's range is simply more iterations of the whole warp -- no per-slot binary search).
B200_KERNEL void __launch_bounds__(256) k_fill_bid(const uint32_t* __restrict__ off1, uint32_t nb, uint32_t* __restrict__ bid1) {
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

// Operand table of a round: one ITEM per ADDITION, written once by k_tree_meta and read (coalesced) by the forward and the backward pass.
// item[e] = (a, b, j, -): operand references -- round 0: the sorted entries (point index | sign << 31), later rounds: positions in the
// previous round's output -- and the output position j.  Slots that only carry a bucket's odd last point over are NOT items: about every
// second bucket has one per round, which is 3 % of the slots in round 0 but 20-33 % in the last rounds (2-4 points per bucket), and a carried
// slot idles its lane for the five multiplications of the warp's additions.  k_tree_meta copies those points itself, so the arithmetic
// kernels run dense: lane utilisation is 32/32 everywhere but in the last warp.  Where an addition's item goes needs no extra scan:
// the additions of the buckets below b number sum (n_r - n_(r+1)) = (off_in[b] - off_in[0]) - off_out[b].
// The table takes the dependent lookups (bid -> offsets -> sorted entry) out of the arithmetic kernels, whose gathers then depend on ONE
// coalesced 16-byte load that is issued an iteration ahead.
// Carried points: carries[k] = (operand reference, output position), k = number of carrying buckets below b = 2*off_out[b] - (off_in[b] - off_in[0]);
// the backward pass copies them after its additions (its CTAs are multiplier-bound, the copies ride along for free; copied by k_tree_meta
// itself they cost 0.2 ms per 2^20-point MSM, latency-bound).
template <bool FIRST>
__global__ void __launch_bounds__(256) k_tree_meta(TreeRound tr, const uint32_t* __restrict__ sorted, uint4* __restrict__ items, uint2* __restrict__ carries) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t in0 = 0; bool has2 = false;
  if (!tree_slot(tr, j, in0, has2, true)) return;
  const uint32_t b = tr.bid[j], oo = tr.off_out[b], adds_below = tr.off_in[b] - tr.off_in[0] - oo;      // off_in may start at the group's position in sorted[]; off_out starts at 0
  const uint32_t ra = FIRST ? __ldg(sorted + in0) : in0;
  if (has2) items[adds_below + (j - oo)] = make_uint4(ra, FIRST ? __ldg(sorted + in0 + 1) : in0 + 1, j, 0u);
  else carries[oo - adds_below] = make_uint2(ra, j);
}

// Round 0 reads the caller's bases (x || y records, 2*n8 bytes apart).  The rounds' own outputs are stored as TWO arrays, all x
// coordinates then (yoff bytes further) all y coordinates, because the forward pass needs only x: its sequential reads halve.
template <class C, bool FIRST>
B200_DI void meta_load_point(Affine<C>& p, const void* __restrict__ src, uint64_t yoff, uint32_t ref) {
  if (FIRST) { affine_load<C>(p, src, ref & 0x7fffffffu); if (ref >> 31) fe_neg<C>(p.y, p.y); }
  else { const char* b = reinterpret_cast<const char*>(src) + (uint64_t)ref * (4 * C::N); fe_load_cg<C>(p.x, b); fe_load_cg<C>(p.y, b + yoff); }
}
// xs (round 0 only, may be null): a copy of the bases' x coordinates alone (k_extract_x), n8 bytes apart.  Every base is gathered once per
// window by the forward pass; 2 * n8 bytes apart the 2^20 x coordinates of BLS12-381 spread over 96 MiB and every gather goes to HBM, n8
// bytes apart they are 48 MiB, which the 126 MB L2 keeps between the windows (no persisting set-aside: carving it out of L2 slowed every other kernel, 2^20: 6.3 -> 8.1 ms).
template <class C, bool FIRST>
B200_DI void meta_load_x(Fe<C::N>& x, const void* __restrict__ src, const void* __restrict__ xs, uint32_t ref) {
  if (FIRST) { if (xs) fe_load<C>(x, reinterpret_cast<const char*>(xs) + (uint64_t)(ref & 0x7fffffffu) * (4 * C::N));
               else fe_load<C>(x, reinterpret_cast<const char*>(src) + (uint64_t)(ref & 0x7fffffffu) * (8 * C::N)); }
  else fe_load_cg<C>(x, reinterpret_cast<const char*>(src) + (uint64_t)ref * (4 * C::N));
}
template <class C>
__global__ void __launch_bounds__(256) k_extract_x(const void* __restrict__ bases, uint32_t n, void* __restrict__ xs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<C::N> x; fe_load<C>(x, reinterpret_cast<const char*>(bases) + (uint64_t)i * (8 * C::N));
  fe_store<C>(reinterpret_cast<char*>(xs) + (uint64_t)i * (4 * C::N), x);
}
template <class C>
B200_DI void soa_store_point(void* __restrict__ dst, uint64_t yoff, uint32_t j, const Affine<C>& p) {
  char* b = reinterpret_cast<char*>(dst) + (uint64_t)j * (4 * C::N); fe_store<C>(b, p.x); fe_store<C>(b + yoff, p.y);
}

// forward: denominators, per-item prefix products, per-thread products.
// The kernel is gather-bound (two scattered 96-byte points per addition in round 0), so the x coordinates of item i+1 are requested
// before the multiplication of item i and the operand references of item i+2 before that (only the x coordinates are needed;
// the y coordinates are fetched in the rare equal-x / zero-x cases).
// One tile (K * BA_THREADS consecutive items, thread t owns items t, t + 128, ...): returns the thread's running product in p.
// nadd = number of items of the round (read from the scan totals on the device: the host only knows an upper bound).
template <class C, bool FIRST>
B200_DI void tree_fwd_tile(Fe<C::N>& p, const uint4* __restrict__ items, uint32_t nadd, const void* __restrict__ src, const void* __restrict__ xs, uint64_t yoff, void* __restrict__ prefix, int K, uint32_t tb) {
  const uint32_t tile = tb * (K * BA_THREADS) + threadIdx.x;
  fe_set_one<C>(p);
  Fe<C::N> x1, x2, nx1, nx2;
  uint4 mc = make_uint4(0, 0, 0, 0), mn = mc;
  if (tile < nadd) mc = items[tile];
  if (K > 1 && tile + BA_THREADS < nadd) mn = items[tile + BA_THREADS];
  if (tile < nadd) { meta_load_x<C, FIRST>(x1, src, xs, mc.x); meta_load_x<C, FIRST>(x2, src, xs, mc.y); }
#pragma unroll 1
  for (int i = 0; i < K; i++) {
    const uint32_t e = tile + i * BA_THREADS;
    uint4 mn2 = make_uint4(0, 0, 0, 0);
    if (i + 2 < K && e + 2 * BA_THREADS < nadd) mn2 = items[e + 2 * BA_THREADS];
    if (i + 1 < K && e + BA_THREADS < nadd) { meta_load_x<C, FIRST>(nx1, src, xs, mn.x); meta_load_x<C, FIRST>(nx2, src, xs, mn.y); }
    if (e < nadd) {
      Fe<C::N> d; int kind = 0;
      fe_sub<C>(d, x2, x1);
      if (fe_is_zero<C>(d) || fe_is_zero<C>(x1) || fe_is_zero<C>(x2)) {        // rare: decide with the full points
        Affine<C> p1, p2;
        meta_load_point<C, FIRST>(p1, src, yoff, mc.x);
        meta_load_point<C, FIRST>(p2, src, yoff, mc.y);
        kind = affine_add_denominator<C>(d, p1, p2);
      }
      if (kind <= 1) {
        fe_store<C>(reinterpret_cast<char*>(prefix) + (uint64_t)e * 4 * C::N, p);
        fe_mul<C>(p, p, d);
      }
    }
    mc = mn; mn = mn2; x1 = nx1; x2 = nx2;
  }
}
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 3 : C::N <= 8 ? B200_FWD_CTAS_N8 : B200_FWD_CTAS_N12) k_tree_fwd(const uint4* __restrict__ items, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                            const void* __restrict__ src, const void* __restrict__ xs, uint64_t yoff, void* __restrict__ prefix, void* __restrict__ prod, int K, uint32_t ntiles) {
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];      // additions of this round = inputs - outputs
 // persistent form: gridDim.x may be smaller than ntiles (leaves SM room for the other lane's latency-bound kernels)
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> p;
  tree_fwd_tile<C, FIRST>(p, items, nadd, src, xs, yoff, prefix, K, tb);
  fe_store<C>(reinterpret_cast<char*>(prod) + (uint64_t)(tb * BA_THREADS + threadIdx.x) * 4 * C::N, p);
 }
}

// backward: consume the inverse q of the thread's product, finish every addition, write the round's output points
template <class C, bool FIRST>
B200_DI void tree_bwd_tile(Fe<C::N>& q, const uint4* __restrict__ items, uint32_t nadd, const void* __restrict__ src, uint64_t yoff, const void* __restrict__ prefix,
                           void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t tb) {
  const uint32_t tile = tb * (K * BA_THREADS) + threadIdx.x;
  uint4 mn = make_uint4(0, 0, 0, 0);
  if (tile + (K - 1) * BA_THREADS < nadd) mn = items[tile + (K - 1) * BA_THREADS];
#pragma unroll 1
  for (int i = K - 1; i >= 0; i--) {
    const uint4 m = mn;
    const uint32_t e = tile + i * BA_THREADS;
    if (i > 0 && e - BA_THREADS < nadd) mn = items[e - BA_THREADS];
    if (e >= nadd) continue;
    Affine<C> p1, p2, r;
    meta_load_point<C, FIRST>(p1, src, yoff, m.x);
    meta_load_point<C, FIRST>(p2, src, yoff, m.y);
    Fe<C::N> d, dinv;
    int kind = affine_add_denominator<C>(d, p1, p2);
    if (kind <= 1) {
      Fe<C::N> pre; fe_load_cg<C>(pre, reinterpret_cast<const char*>(prefix) + (uint64_t)e * 4 * C::N);
      // the two multiplications of the inverse-sharing step are independent: their rows are interleaved (fe_mul2), which doubles the
      // work between dependent carry-chain instructions (measured: k_tree_bwd 3.46 -> 3.39 ms at 2^20, profiles/README.md r2)
      if constexpr (C::EXT == 1) { Fe<C::N> qn; fe_mul2<C>(dinv, q, pre, qn, q, d); q = qn; }
      else { fe_mul<C>(dinv, q, pre); fe_mul<C>(q, q, d); }
    }
    affine_add_finish<C>(r, p1, p2, dinv, kind);
    soa_store_point<C>(pout, yoff_out, m.z, r);
  }
}
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 2 : C::N <= 8 ? B200_BWD_CTAS_N8 : 4) k_tree_bwd(const uint4* __restrict__ items, const uint2* __restrict__ carries, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                         const void* __restrict__ src, uint64_t yoff,
                                                         const void* __restrict__ prefix, const void* __restrict__ inv,
                                                         void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles) {
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];      // additions of this round = inputs - outputs
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> q;
  fe_load_cg<C>(q, reinterpret_cast<const char*>(inv) + (uint64_t)(tb * BA_THREADS + threadIdx.x) * 4 * C::N);
  tree_bwd_tile<C, FIRST>(q, items, nadd, src, yoff, prefix, pout, yoff_out, K, tb);
 }
 // the round's carried points (a bucket's odd last input): plain copies into their output slots (round 0: with the digit's sign applied)
 const uint32_t ncar = off_out[nb] - nadd;
 for (uint32_t c = blockIdx.x * BA_THREADS + threadIdx.x; c < ncar; c += gridDim.x * BA_THREADS) {
  const uint2 cr = carries[c];
  Affine<C> p; meta_load_point<C, FIRST>(p, src, yoff, cr.x); soa_store_point<C>(pout, yoff_out, cr.y, p);
 }
}

// ---- one tree round as ONE kernel: every CTA inverts the product of ITS OWN tile ------------------------------------------------------
// MEASURED AND NOT ADOPTED (round 2, profiles/README.md "r2 late"): 2^20 BLS12-381 8.2-9.8 ms against 6.29 (four lanes), 64 x 2^18 in a batch 2.24 ms per MSM
// against 1.70, BN254 2^20 4.17 against 3.44 -- the CTAs of a launch run in step (all in their memory-bound forward phase, then all waiting for
// their inverting thread, then all in the multiplier-bound backward phase), so the phases add up instead of overlapping, which separate
// forward / backward launches on several lanes avoid.  Kept for -DB200_EXPERIMENTS builds only (option "fused_round").
#if defined(B200_EXPERIMENTS)
// f1m_batchInverse (build_batchinverse.js:4-140) needs one field inversion per batch, and the reference runs one batch per level of its
// addition chains.  The grid-wide form above does the same -- one inversion per round -- and pays for it with a chain of launches
// (forward pass, product-tree levels, root, levels back down, backward pass) whose latency-bound middle is ~0.2 ms per round per lane.
// Here the batch is a TILE (K * BA_THREADS additions): a CTA runs the forward pass of its tile, reduces the 128 per-thread products by a
// binary tree in shared memory, thread 0 inverts the tile's root (Pornin's binary GCD, ~50 us on one thread, ~1 % of the tile's instructions),
// the tree is walked back down and the backward pass follows -- no grid-wide dependency, no product-tree launches, one launch per round.
// The other CTAs resident on the SM (different phases of their own tiles) keep the multiplier busy while one thread inverts.
// Cost per addition: the same 6 multiplications + 3 / K for the block tree (the grid-wide product tree's first level costs the same 3 / K).
// ---- block-level product tree (the first level of the round's batch inversion lives INSIDE the tree kernels) ---------------------------
// MEASURED AND NOT ADOPTED (round 2; k_tree_fwd_bt / k_tree_bwd_bt, -DB200_EXPERIMENTS builds only, option "block_tree"): 2^16 1.476 -> 1.457 ms,
// 2^18 2.63 -> 2.61, but 2^20 6.25 -> 6.55 and 2^22 20.5 -> 21.6, batches 1.70 -> 1.77 ms per MSM: the down-sweep at the head of every backward tile is a
// latency-bound phase that all CTAs of the launch enter together (k_tree_bwd round 0: 1.53 -> 1.75 ms), which costs more than the product-tree launches it removes.
// The 128 per-thread products of a tile are reduced by a binary tree in shared memory at the end of the forward pass, so ONE value per
// tile goes up to the grid-wide product tree (a round of 2 M additions then needs no product-tree launch at all: <= 1024 tile roots go
// straight to k_inv_root); the tree's nodes are kept (256 elements per tile) and the backward pass walks them down from the tile root's
// inverse before its additions.  Same multiplications as a K-ary level of the grid-wide tree (3 per value), two to four launches and
// their latency-bound tails fewer per round.  Heap order: node 1 = root, leaves at [BA_THREADS, 2 * BA_THREADS).
template <class C>
B200_DI void block_upsweep(const Fe<C::N>& p, uint32_t* __restrict__ tree, void* __restrict__ gtree) {      // gtree: this tile's 2 * BA_THREADS elements in global memory
  constexpr int N = C::N;
  const uint32_t t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < N; k++) tree[(BA_THREADS + t) * N + k] = p.l[k];
  fe_store<C>(reinterpret_cast<char*>(gtree) + (uint64_t)(BA_THREADS + t) * 4 * N, p);
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
      fe_store<C>(reinterpret_cast<char*>(gtree) + (uint64_t)node * 4 * N, c);
    }
    __syncthreads();
  }
}
// tree[1] holds the inverse of the root on entry (written before the call, barrier included here); on exit q = inverse of the thread's own product
template <class C>
B200_DI void block_downsweep(Fe<C::N>& q, uint32_t* __restrict__ tree) {
  constexpr int N = C::N;
  const uint32_t t = threadIdx.x;
  __syncthreads();
#pragma unroll 1
  for (uint32_t width = 1; width < BA_THREADS; width <<= 1) {          // one thread per CHILD: inv(child) = inv(parent) * sibling
    Fe<N> ip, sib, r; const uint32_t child = 2 * width + t;
    if (t < 2 * width) {
#pragma unroll
      for (int k = 0; k < N; k++) { ip.l[k] = tree[(child >> 1) * N + k]; sib.l[k] = tree[(child ^ 1) * N + k]; }
    }
    __syncthreads();
    if (t < 2 * width) {
      fe_mul<C>(r, ip, sib);
#pragma unroll
      for (int k = 0; k < N; k++) tree[child * N + k] = r.l[k];
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < N; k++) q.l[k] = tree[(BA_THREADS + t) * N + k];
}

template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 3 : 6) k_tree_fwd_bt(const uint4* __restrict__ items, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                            const void* __restrict__ src, const void* __restrict__ xs, uint64_t yoff, void* __restrict__ prefix, void* __restrict__ prod, int K, uint32_t ntiles, void* __restrict__ gtree) {
 __shared__ __align__(16) uint32_t tree[2 * BA_THREADS * C::N];
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];      // additions of this round = inputs - outputs
 // persistent form: gridDim.x may be smaller than ntiles (leaves SM room for the other lane's latency-bound kernels)
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> p;
  tree_fwd_tile<C, FIRST>(p, items, nadd, src, xs, yoff, prefix, K, tb);
  {      // one value per TILE goes up (tiles beyond the round's additions contribute 1)
    block_upsweep<C>(p, tree, reinterpret_cast<char*>(gtree) + (uint64_t)tb * (2 * BA_THREADS) * 4 * C::N);
    if (threadIdx.x == 0) { Fe<C::N> r;
#pragma unroll
      for (int k = 0; k < C::N; k++) r.l[k] = tree[C::N + k];
      fe_store<C>(reinterpret_cast<char*>(prod) + (uint64_t)tb * 4 * C::N, r); }
  }
 }
}


template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 2 : 4) k_tree_bwd_bt(const uint4* __restrict__ items, const uint2* __restrict__ carries, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                         const void* __restrict__ src, uint64_t yoff,
                                                         const void* __restrict__ prefix, const void* __restrict__ inv,
                                                         void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles, const void* __restrict__ gtree) {
 __shared__ __align__(16) uint32_t tree[2 * BA_THREADS * C::N];
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];      // additions of this round = inputs - outputs
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  Fe<C::N> q;
  {      // the tile's tree (written by the forward pass) with the root replaced by its inverse, walked down to the threads
    const char* gt = reinterpret_cast<const char*>(gtree) + (uint64_t)tb * (2 * BA_THREADS) * 4 * C::N;
    const uint32_t t = threadIdx.x;
    Fe<C::N> a, b;
    fe_load_cg<C>(a, t <= 1 ? reinterpret_cast<const char*>(inv) + (uint64_t)tb * 4 * C::N : gt + (uint64_t)t * 4 * C::N);      // node 0 is unused; node 1 <- inverse of the root
    fe_load_cg<C>(b, gt + (uint64_t)(BA_THREADS + t) * 4 * C::N);
    __syncthreads();                                                        // the previous tile's down-sweep has been read by every thread
#pragma unroll
    for (int k = 0; k < C::N; k++) { tree[t * C::N + k] = a.l[k]; tree[(BA_THREADS + t) * C::N + k] = b.l[k]; }
    block_downsweep<C>(q, tree);
  }
  tree_bwd_tile<C, FIRST>(q, items, nadd, src, yoff, prefix, pout, yoff_out, K, tb);
 }
 // the round's carried points (a bucket's odd last input): plain copies into their output slots (round 0: with the digit's sign applied)
 const uint32_t ncar = off_out[nb] - nadd;
 for (uint32_t c = blockIdx.x * BA_THREADS + threadIdx.x; c < ncar; c += gridDim.x * BA_THREADS) {
  const uint2 cr = carries[c];
  Affine<C> p; meta_load_point<C, FIRST>(p, src, yoff, cr.x); soa_store_point<C>(pout, yoff_out, cr.y, p);
 }
}


template <class C>
B200_DI void block_invert(Fe<C::N>& q, const Fe<C::N>& p, uint32_t* __restrict__ tree) {     // tree: 2 * BA_THREADS elements, heap order, leaves at [BA_THREADS, 2 * BA_THREADS)
  constexpr int N = C::N;
  const uint32_t t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < N; k++) tree[(BA_THREADS + t) * N + k] = p.l[k];
  __syncthreads();
#pragma unroll 1
  for (uint32_t width = BA_THREADS / 2; width >= 1; width >>= 1) {     // up-sweep: node = left * right
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
  if (t == 0) {
    Fe<N> r, ri;
#pragma unroll
    for (int k = 0; k < N; k++) r.l[k] = tree[N + k];
    fe_inv_fast<C>(ri, r);
#pragma unroll
    for (int k = 0; k < N; k++) tree[N + k] = ri.l[k];
  }
  __syncthreads();
#pragma unroll 1
  for (uint32_t width = 1; width < BA_THREADS; width <<= 1) {          // down-sweep, one thread per CHILD: inv(child) = inv(node) * sibling
    Fe<N> ip, sib, r; const uint32_t child = 2 * width + t;
    if (t < 2 * width) {
#pragma unroll
      for (int k = 0; k < N; k++) { ip.l[k] = tree[(child >> 1) * N + k]; sib.l[k] = tree[(child ^ 1) * N + k]; }
    }
    __syncthreads();
    if (t < 2 * width) {
      fe_mul<C>(r, ip, sib);
#pragma unroll
      for (int k = 0; k < N; k++) tree[child * N + k] = r.l[k];
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < N; k++) q.l[k] = tree[(BA_THREADS + t) * N + k];
}
template <class C, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS, C::N > 12 ? 2 : 4) k_tree_round(const uint4* __restrict__ items, const uint2* __restrict__ carries, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                           const void* __restrict__ src, uint64_t yoff, void* __restrict__ prefix,
                                                           void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles) {
 __shared__ __align__(16) uint32_t tree[2 * BA_THREADS * C::N];
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];      // additions of this round = inputs - outputs
 for (uint32_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
  if (tb * (uint32_t)(K * BA_THREADS) >= nadd) break;             // ntiles comes from the host's upper bound of nadd
  Fe<C::N> p, q;
  tree_fwd_tile<C, FIRST>(p, items, nadd, src, nullptr, yoff, prefix, K, tb);
  block_invert<C>(q, p, tree);
  tree_bwd_tile<C, FIRST>(q, items, nadd, src, yoff, prefix, pout, yoff_out, K, tb);
 }
 const uint32_t ncar = off_out[nb] - nadd;
 for (uint32_t c = blockIdx.x * BA_THREADS + threadIdx.x; c < ncar; c += gridDim.x * BA_THREADS) {
  const uint2 cr = carries[c];
  Affine<C> p; meta_load_point<C, FIRST>(p, src, yoff, cr.x); soa_store_point<C>(pout, yoff_out, cr.y, p);
 }
}

// ---- the same round, software-pipelined inside the CTA (option fused_round = 2) ----------------------------------------------------------
// Four compute warps + ONE inverting warp per CTA, two product-tree buffers: the compute warps run the forward phase of tile i+1, hand its
// root to the inverting warp and go on to the backward phase of tile i, whose root the inverting warp finished meanwhile -- the 50 us of
// the inversion sit under ~100-250 us of multiplications instead of stalling the CTA.  Named barriers: 1 = the compute warps among
// themselves, 2 + b = "root of buffer b is ready" (compute arrives, inverter waits), 4 + b = "its inverse is ready" (the other way round).
// (copies of tree_fwd_tile / tree_bwd_tile that take the index of the tile's first item: the de-phased form below cuts a CTA's first tile in two)
template <class C, bool FIRST>
B200_DI void tree_fwd_tile_b(Fe<C::N>& p, const uint4* __restrict__ items, uint32_t nadd, const void* __restrict__ src, const void* __restrict__ xs, uint64_t yoff, void* __restrict__ prefix, int K, uint32_t tb) {
  const uint32_t tile = tb + threadIdx.x;      // tb = index of the tile's first item
  fe_set_one<C>(p);
  Fe<C::N> x1, x2, nx1, nx2;
  uint4 mc = make_uint4(0, 0, 0, 0), mn = mc;
  if (tile < nadd) mc = items[tile];
  if (K > 1 && tile + BA_THREADS < nadd) mn = items[tile + BA_THREADS];
  if (tile < nadd) { meta_load_x<C, FIRST>(x1, src, xs, mc.x); meta_load_x<C, FIRST>(x2, src, xs, mc.y); }
#pragma unroll 1
  for (int i = 0; i < K; i++) {
    const uint32_t e = tile + i * BA_THREADS;
    uint4 mn2 = make_uint4(0, 0, 0, 0);
    if (i + 2 < K && e + 2 * BA_THREADS < nadd) mn2 = items[e + 2 * BA_THREADS];
    if (i + 1 < K && e + BA_THREADS < nadd) { meta_load_x<C, FIRST>(nx1, src, xs, mn.x); meta_load_x<C, FIRST>(nx2, src, xs, mn.y); }
    if (e < nadd) {
      Fe<C::N> d; int kind = 0;
      fe_sub<C>(d, x2, x1);
      if (fe_is_zero<C>(d) || fe_is_zero<C>(x1) || fe_is_zero<C>(x2)) {        // rare: decide with the full points
        Affine<C> p1, p2;
        meta_load_point<C, FIRST>(p1, src, yoff, mc.x);
        meta_load_point<C, FIRST>(p2, src, yoff, mc.y);
        kind = affine_add_denominator<C>(d, p1, p2);
      }
      if (kind <= 1) {
        fe_store<C>(reinterpret_cast<char*>(prefix) + (uint64_t)e * 4 * C::N, p);
        fe_mul<C>(p, p, d);
      }
    }
    mc = mn; mn = mn2; x1 = nx1; x2 = nx2;
  }
}
template <class C, bool FIRST>
B200_DI void tree_bwd_tile_b(Fe<C::N>& q, const uint4* __restrict__ items, uint32_t nadd, const void* __restrict__ src, uint64_t yoff, const void* __restrict__ prefix,
                           void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t tb) {
  const uint32_t tile = tb + threadIdx.x;      // tb = index of the tile's first item
  uint4 mn = make_uint4(0, 0, 0, 0);
  if (tile + (K - 1) * BA_THREADS < nadd) mn = items[tile + (K - 1) * BA_THREADS];
#pragma unroll 1
  for (int i = K - 1; i >= 0; i--) {
    const uint4 m = mn;
    const uint32_t e = tile + i * BA_THREADS;
    if (i > 0 && e - BA_THREADS < nadd) mn = items[e - BA_THREADS];
    if (e >= nadd) continue;
    Affine<C> p1, p2, r;
    meta_load_point<C, FIRST>(p1, src, yoff, m.x);
    meta_load_point<C, FIRST>(p2, src, yoff, m.y);
    Fe<C::N> d, dinv;
    int kind = affine_add_denominator<C>(d, p1, p2);
    if (kind <= 1) {
      Fe<C::N> pre; fe_load_cg<C>(pre, reinterpret_cast<const char*>(prefix) + (uint64_t)e * 4 * C::N);
      // the two multiplications of the inverse-sharing step are independent: their rows are interleaved (fe_mul2), which doubles the
      // work between dependent carry-chain instructions (measured: k_tree_bwd 3.46 -> 3.39 ms at 2^20, profiles/README.md r2)
      if constexpr (C::EXT == 1) { Fe<C::N> qn; fe_mul2<C>(dinv, q, pre, qn, q, d); q = qn; }
      else { fe_mul<C>(dinv, q, pre); fe_mul<C>(q, q, d); }
    }
    affine_add_finish<C>(r, p1, p2, dinv, kind);
    soa_store_point<C>(pout, yoff_out, m.z, r);
  }
}
B200_DI void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
B200_DI void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
constexpr int RP_THREADS = BA_THREADS + 32;
template <class C>
B200_DI void pipe_upsweep(const Fe<C::N>& p, uint32_t* __restrict__ tree) {
  constexpr int N = C::N;
  const uint32_t t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < N; k++) tree[(BA_THREADS + t) * N + k] = p.l[k];
  bar_sync_n(1, BA_THREADS);
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
    bar_sync_n(1, BA_THREADS);
  }
}
template <class C>
B200_DI void pipe_downsweep(Fe<C::N>& q, uint32_t* __restrict__ tree) {
  constexpr int N = C::N;
  const uint32_t t = threadIdx.x;
#pragma unroll 1
  for (uint32_t width = 1; width < BA_THREADS; width <<= 1) {
    Fe<N> ip, sib, r; const uint32_t child = 2 * width + t;
    if (t < 2 * width) {
#pragma unroll
      for (int k = 0; k < N; k++) { ip.l[k] = tree[(child >> 1) * N + k]; sib.l[k] = tree[(child ^ 1) * N + k]; }
    }
    bar_sync_n(1, BA_THREADS);
    if (t < 2 * width) {
      fe_mul<C>(r, ip, sib);
#pragma unroll
      for (int k = 0; k < N; k++) tree[child * N + k] = r.l[k];
    }
    bar_sync_n(1, BA_THREADS);
  }
#pragma unroll
  for (int k = 0; k < N; k++) q.l[k] = tree[(BA_THREADS + t) * N + k];
}
template <class C, bool FIRST>
__global__ void __launch_bounds__(RP_THREADS, 3) k_tree_round_pipe(const uint4* __restrict__ items, const uint2* __restrict__ carries, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                           const void* __restrict__ src, const void* __restrict__ xs, uint64_t yoff, void* __restrict__ prefix,
                                                           void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles) {
 constexpr int N = C::N;
 __shared__ __align__(16) uint32_t tree[2][2 * BA_THREADS * N];
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];
 const uint32_t tile_sz = (uint32_t)K * BA_THREADS;
 const uint32_t nt = min(ntiles, (nadd + tile_sz - 1) / tile_sz);
 const uint32_t T = blockIdx.x < nt ? (nt - 1 - blockIdx.x) / gridDim.x + 1 : 0;      // tiles blockIdx.x, blockIdx.x + gridDim.x, ...
 if (threadIdx.x >= BA_THREADS) {                                                     // the inverting warp
  for (uint32_t k = 0; k < T; k++) {
   const uint32_t b = k & 1;
   bar_sync_n(2 + b, RP_THREADS);
   if (threadIdx.x == BA_THREADS) {
    Fe<N> r, ri;
#pragma unroll
    for (int j = 0; j < N; j++) r.l[j] = tree[b][N + j];
    fe_inv_fast<C>(ri, r);
#pragma unroll
    for (int j = 0; j < N; j++) tree[b][N + j] = ri.l[j];
   }
   __syncwarp();
   bar_arrive_n(4 + b, RP_THREADS);
  }
  return;
 }
 if (T) { Fe<N> p; tree_fwd_tile<C, FIRST>(p, items, nadd, src, xs, yoff, prefix, K, blockIdx.x); pipe_upsweep<C>(p, tree[0]); bar_arrive_n(2, RP_THREADS); }
 for (uint32_t k = 0; k < T; k++) {
  const uint32_t b = k & 1, tb = blockIdx.x + k * gridDim.x;
  if (k + 1 < T) { Fe<N> p; tree_fwd_tile<C, FIRST>(p, items, nadd, src, xs, yoff, prefix, K, tb + gridDim.x); pipe_upsweep<C>(p, tree[b ^ 1]); bar_arrive_n(2 + (b ^ 1), RP_THREADS); }
  bar_sync_n(4 + b, RP_THREADS);
  Fe<N> q;
  pipe_downsweep<C>(q, tree[b]);
  tree_bwd_tile<C, FIRST>(q, items, nadd, src, yoff, prefix, pout, yoff_out, K, tb);
 }
 const uint32_t ncar = off_out[nb] - nadd;
 for (uint32_t c = blockIdx.x * BA_THREADS + threadIdx.x; c < ncar; c += gridDim.x * BA_THREADS) {
  const uint2 cr = carries[c];
  Affine<C> p; meta_load_point<C, FIRST>(p, src, yoff, cr.x); soa_store_point<C>(pout, yoff_out, cr.y, p);
 }
}
// De-phased form (fused_round = 3): the CTAs that share an SM (blockIdx.x / phase_div = 0, 1, 2 for a grid of three CTAs per SM) cut their FIRST tile
// in two pieces of j/3 and (3 - j)/3 of its length, so that from then on they sit a third of a tile period apart: while one gathers (forward
// phase), the others multiply.
template <class C, bool FIRST>
__global__ void __launch_bounds__(RP_THREADS, 3) k_tree_round_pipe3(const uint4* __restrict__ items, const uint2* __restrict__ carries, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t nb,
                                                           const void* __restrict__ src, const void* __restrict__ xs, uint64_t yoff, void* __restrict__ prefix,
                                                           void* __restrict__ pout, uint64_t yoff_out, int K, uint32_t ntiles, uint32_t phase_div) {
 constexpr int N = C::N;
 __shared__ __align__(16) uint32_t tree[2][2 * BA_THREADS * N];
 const uint32_t nadd = off_in[nb] - off_in[0] - off_out[nb];
 const uint32_t tile_sz = (uint32_t)K * BA_THREADS;
 const uint32_t nt = min(ntiles, (nadd + tile_sz - 1) / tile_sz);
 const uint32_t T = blockIdx.x < nt ? (nt - 1 - blockIdx.x) / gridDim.x + 1 : 0;
 const uint32_t j = (blockIdx.x / phase_div) % 3, K0 = (uint32_t)K * j / 3;                 // first piece: K0 items per thread (0: the tile is not cut)
 const uint32_t U = T + ((K0 && T) ? 1 : 0);                                                // units = tiles + 1 when the first tile is cut
 auto unit = [&](uint32_t u, uint32_t& base, int& Ku) {
  if (K0 == 0) { base = (blockIdx.x + u * gridDim.x) * tile_sz; Ku = K; return; }
  if (u == 0) { base = blockIdx.x * tile_sz; Ku = (int)K0; return; }
  if (u == 1) { base = blockIdx.x * tile_sz + K0 * BA_THREADS; Ku = K - (int)K0; return; }
  base = (blockIdx.x + (u - 1) * gridDim.x) * tile_sz; Ku = K;
 };
 if (threadIdx.x >= BA_THREADS) {
  for (uint32_t k = 0; k < U; k++) {
   const uint32_t b = k & 1;
   bar_sync_n(2 + b, RP_THREADS);
   if (threadIdx.x == BA_THREADS) {
    Fe<N> r, ri;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = tree[b][N + i];
    fe_inv_fast<C>(ri, r);
#pragma unroll
    for (int i = 0; i < N; i++) tree[b][N + i] = ri.l[i];
   }
   __syncwarp();
   bar_arrive_n(4 + b, RP_THREADS);
  }
  return;
 }
 uint32_t base; int Ku;
 if (U) { Fe<N> p; unit(0, base, Ku); tree_fwd_tile_b<C, FIRST>(p, items, nadd, src, xs, yoff, prefix, Ku, base); pipe_upsweep<C>(p, tree[0]); bar_arrive_n(2, RP_THREADS); }
 for (uint32_t k = 0; k < U; k++) {
  const uint32_t b = k & 1;
  if (k + 1 < U) { Fe<N> p; unit(k + 1, base, Ku); tree_fwd_tile_b<C, FIRST>(p, items, nadd, src, xs, yoff, prefix, Ku, base); pipe_upsweep<C>(p, tree[b ^ 1]); bar_arrive_n(2 + (b ^ 1), RP_THREADS); }
  bar_sync_n(4 + b, RP_THREADS);
  Fe<N> q;
  pipe_downsweep<C>(q, tree[b]);
  unit(k, base, Ku);
  tree_bwd_tile_b<C, FIRST>(q, items, nadd, src, yoff, prefix, pout, yoff_out, Ku, base);
 }
 const uint32_t ncar = off_out[nb] - nadd;
 for (uint32_t c = blockIdx.x * BA_THREADS + threadIdx.x; c < ncar; c += gridDim.x * BA_THREADS) {
  const uint2 cr = carries[c];
  Affine<C> p; meta_load_point<C, FIRST>(p, src, yoff, cr.x); soa_store_point<C>(pout, yoff_out, cr.y, p);
 }
}
#endif  // B200_EXPERIMENTS

// Measured alternatives that are no longer in the tree (profiles/README.md has their numbers): a backward pass that stages the next slot's operands
// in shared memory with cp.async (0-8 % slower), a register-lean backward pass that reloads operands instead of keeping them (5 CTAs/SM: equal
// or slower), and the whole round as ONE cooperative launch with in-kernel wave-wise batch inversion (2^20: 8.2-10.6 ms against 6.5).

// Explicit instantiations of the two hot kernels live in tree.cu (one object per curve, compiled without -split-compile: see there);
// every other translation unit sees them as `extern template`.
#define B200_TREE_INSTANTIATE(KW, C) \
  KW __global__ void k_tree_fwd<C, true>(const uint4* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__, uint32_t, const void* __restrict__, const void* __restrict__, uint64_t, void* __restrict__, void* __restrict__, int, uint32_t); \
  KW __global__ void k_tree_fwd<C, false>(const uint4* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__, uint32_t, const void* __restrict__, const void* __restrict__, uint64_t, void* __restrict__, void* __restrict__, int, uint32_t); \
  KW __global__ void k_tree_bwd<C, true>(const uint4* __restrict__, const uint2* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__, uint32_t, const void* __restrict__, uint64_t, const void* __restrict__, const void* __restrict__, void* __restrict__, uint64_t, int, uint32_t); \
  KW __global__ void k_tree_bwd<C, false>(const uint4* __restrict__, const uint2* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__, uint32_t, const void* __restrict__, uint64_t, const void* __restrict__, const void* __restrict__, void* __restrict__, uint64_t, int, uint32_t);
#if !defined(TREE_CURVE)
B200_TREE_INSTANTIATE(extern template, BLS12_381)
B200_TREE_INSTANTIATE(extern template, BN254)
#if !defined(B200_NO_G2)
B200_TREE_INSTANTIATE(extern template, Fq2<BLS12_381>)
B200_TREE_INSTANTIATE(extern template, Fq2<BN254>)
#endif
#endif

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
  for (uint32_t width = 1; width < BA_ROOT_THREADS; width <<= 1) {          // down-sweep, one thread per CHILD: inv(child) = inv(parent) * sibling -- one
    Fe<N> ip, sib, r; const uint32_t child = 2 * width + t;                  // multiplication deep per level instead of two (the kernel is one dependent chain)
    if (t < 2 * width) {
#pragma unroll
      for (int k = 0; k < N; k++) { ip.l[k] = tree[(child >> 1) * N + k]; sib.l[k] = tree[(child ^ 1) * N + k]; }
    }
    __syncthreads();
    if (t < 2 * width) {
      fe_mul<C>(r, ip, sib);
#pragma unroll
      for (int k = 0; k < N; k++) tree[child * N + k] = r.l[k];
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
