// msm_kernels.cuh -- the Pippenger pipeline as sm_100a kernels.
//
// Path replaced: wasmcurves/src/build_multiexp.js:25-461 (getChunk, _chunk,
// _reduceTable, multiexp) and the schedule / bucket phases of the Manta opt path,
// wasmcurves/src/build_multiexp_opt.js:175-347 (computeSchedule), :364-633 (organizeBuckets),
// :1336-1585 (reduceBuckets), :1597-1706 (reduceBucketsToSinglePoint), :1710-1746 (accumulateAcrossChunks).
//
// Pipeline for one MSM over the scalar bit range [bit0, bit0 + nbits):
//   1. k_digits<COUNT>     signed-digit recoding of every scalar (window width c, digits in [-2^(c-1), 2^(c-1)]),
//                          histogram of (window, |digit|) with global atomics                       [HBM / atomics]
//   2. k_scan_*            exclusive scan of the W*B bucket counters -> segment offsets            [HBM]
//   3. k_digits<SCATTER>   same recoding, each (point, sign) written into its bucket segment        [HBM / atomics]
//   4. accumulate          per-bucket sums (see accumulate kernels)                                  [IMAD]
//   5. k_fold              log-depth in-place folding of every window's bucket array:
//                          after p = log2(B) levels  sum_b b*T[b-1] = T[0] + sum_j 2^j T[2^j]       [IMAD]
//   6. k_window_sums, k_horner  per-window totals, then result = sum_w 2^(c*w) R_w                   [latency]
#pragma once
#include "ec.cuh"

namespace b200 {

struct MsmPlan {
  uint32_t n;          // points
  uint32_t c;          // widest window in bits (= c0 + 1 if rem > 0 else c0)
  uint32_t c0, rem;    // the nbits are split into Wd windows whose widths differ by at most one bit: windows 0..rem-1 are
                       // c0 + 1 bits wide, windows rem..Wd-1 are c0 bits wide (bit offset of window w: w*c0 + min(w, rem))
  uint32_t Wd;         // digit windows: 0..Wd-2 are signed (digit in [-2^(cw-1), 2^(cw-1)]), the last one is unsigned and
                       // absorbs the final carry (digit in [0, 2^c0])
  uint32_t W;          // bucket-array slots of B buckets each: Wd, or Wd + 1 when the last window needs 2B buckets (rem == 0)
  uint32_t B;          // buckets per slot = 2^(c-1)   (bucket index = |digit| - 1)
  uint32_t nbits;      // scalar bits processed
  uint32_t logB;       // c - 1
  uint32_t pre_stride; // 0: one bucket array per window, points = bases[i].  > 0 (precomputed window tables, see k_table_double):
                       // ONE bucket array for all windows, digit w of scalar i selects table[w * pre_stride + i] = 2^(bit offset of w) * P_i
};

// ------------------------------------------------------------------ scalars
// Canonical scalar layout inside the engine: 8 u32 words (256 bit) little-endian per scalar, already
// shifted so that bit 0 is the first processed bit and masked to nbits.  For the common case
// (scalar_size == 32, bit0 == 0, nbits == 256) the caller's buffer is used in place.
B200_KERNEL void k_canon_scalars(const uint8_t* __restrict__ in, uint32_t scalar_size, uint32_t n,
                                uint32_t bit0, uint32_t nbits, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* s = in + (uint64_t)i * scalar_size;
  uint32_t w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = 0;
  // gather the bytes covering [bit0, bit0 + nbits)
  uint32_t byte0 = bit0 >> 3, sh = bit0 & 7;
  for (uint32_t k = 0; k < 33; k++) {
    uint32_t src = byte0 + k;
    uint32_t v = (src < scalar_size) ? s[src] : 0u;
    // place byte k at bit position 8k - sh (may straddle words)
    int bitpos = (int)(8 * k) - (int)sh;
    if (bitpos >= 0) {
      if (bitpos < 256) { w[bitpos >> 5] |= v << (bitpos & 31); if ((bitpos & 31) > 24) w[(bitpos >> 5) + 1] |= v >> (32 - (bitpos & 31)); }
    } else {
      w[0] |= v >> sh;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t lo = 32 * k;
    uint32_t m = (nbits >= lo + 32) ? 0xffffffffu : (nbits <= lo ? 0u : ((1u << (nbits - lo)) - 1u));
    out[(uint64_t)i * 8 + k] = w[k] & m;
  }
}

// Digit recoding.  Windows below the last one use signed digits with a carry into the next window; the last window keeps
// its digit unsigned (it may reach 2^c0 and then indexes into the extra slot), so no carry ever leaves the scalar and no
// "carry-only" window exists.  A non-zero digit d of window w lands in global bucket w*B + |d| - 1 with its sign.
// Digits of 8 consecutive windows at a time: the recoding is a serial carry chain, but the 8 atomics that follow are independent,
// so they are issued back to back instead of one L2 round trip per window.
// Two passes over the scalars (computeSchedule + organizeBuckets, build_multiexp_opt.js:175-633):
//   COUNT   : histogram of (window, |digit|); the value each atomicAdd returns is the pair's RANK inside its bucket and is kept
//             (ranks[w * n + i], coalesced per window), so that
//   SCATTER : needs no second round of atomics: position = offsets[bucket] + rank, one L2-resident table lookup per pair.
// [wlo, whi): the windows whose pairs this launch emits (the recoding still walks the windows below wlo for the carry).  The window
// slots are sorted group by group on the lanes' own streams, so that a lane's tree starts as soon as ITS windows are sorted.
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ counters,
                                                const uint32_t* __restrict__ offsets, uint32_t* __restrict__ ranks, uint32_t* __restrict__ sorted,
                                                uint32_t wlo, uint32_t whi) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pl.n) return;
  const uint32_t* s = scalars + (uint64_t)i * 8;
  uint32_t sw[8];
  { const uint4 a = __ldg(reinterpret_cast<const uint4*>(s)), b = __ldg(reinterpret_cast<const uint4*>(s) + 1);
    sw[0] = a.x; sw[1] = a.y; sw[2] = a.z; sw[3] = a.w; sw[4] = b.x; sw[5] = b.y; sw[6] = b.z; sw[7] = b.w; }
  uint32_t carry = 0;
  const uint32_t wend = min(pl.Wd, whi);
  for (uint32_t w0 = 0; w0 < wend; w0 += 8) {
    uint32_t gb[8], val[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t w = w0 + u;
      gb[u] = 0xffffffffu; val[u] = i;
      if (w < pl.Wd) {
        const uint32_t cw = pl.c0 + (w < pl.rem ? 1u : 0u);
        const uint32_t bit = w * pl.c0 + min(w, pl.rem), k = bit >> 5, r = bit & 31;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { lo = (k == (uint32_t)q) ? sw[q] : lo; hi = (k + 1 == (uint32_t)q) ? sw[q] : hi; }
        const uint32_t raw = __funnelshift_r(lo, hi, r) & ((1u << cw) - 1u);
        const uint32_t d = raw + carry;
        const uint32_t slot0 = pl.pre_stride ? 0u : w * pl.B;
        if (pl.pre_stride) val[u] = i + w * pl.pre_stride;
        const bool emit = w >= wlo && w < whi;
        if (w + 1 == pl.Wd) { if (d && emit) gb[u] = slot0 + d - 1; }
        else {
          carry = d > (1u << (cw - 1));
          const uint32_t mag = carry ? ((1u << cw) - d) : d;
          if (mag && emit) { gb[u] = slot0 + mag - 1; val[u] |= carry << 31; }
        }
      }
    }
    if (!SCATTER) {
      uint32_t rk[8];
#pragma unroll
      for (int u = 0; u < 8; u++) if (gb[u] != 0xffffffffu) rk[u] = atomicAdd(&counters[gb[u]], 1u);
#pragma unroll
      for (int u = 0; u < 8; u++) if (gb[u] != 0xffffffffu) ranks[(uint64_t)(w0 + u) * pl.n + i] = rk[u];
    } else {
      uint32_t pos[8];
#pragma unroll
      for (int u = 0; u < 8; u++) if (gb[u] != 0xffffffffu) pos[u] = __ldg(offsets + gb[u]) + ranks[(uint64_t)(w0 + u) * pl.n + i];
#pragma unroll
      for (int u = 0; u < 8; u++) if (gb[u] != 0xffffffffu) sorted[pos[u]] = val[u];
    }
  }
}

// ------------------------------------------------------------------ exclusive scan (u32), three phases
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 16, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = (lane < blockDim.x / 32) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    warp_sums[lane] = s;
  }
  __syncthreads();
  uint32_t base = wid ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[blockDim.x / 32 - 1];
  __syncthreads();
  return base + x - v;
}

B200_KERNEL void k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ tile_sums) {
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single block: scan tile sums in place (exclusive), total appended at tile_sums[ntiles]
B200_KERNEL void k_scan_sums(uint32_t* __restrict__ tile_sums, uint32_t ntiles, uint32_t base) {      // base: position of the first element's segment (a group's region of sorted[])
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = base;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = (i < ntiles) ? tile_sums[i] : 0, total;
    uint32_t ex = block_exclusive_scan(v, &total);
    uint32_t c = carry_s;
    if (i < ntiles) tile_sums[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[ntiles] = carry_s;
}
// add tile offsets; writes offsets[n] = grand total; optionally copies the offsets into a cursor array
B200_KERNEL void k_scan_apply(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                             uint32_t* __restrict__ cursors) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { uint32_t v = out[i] + tile_sums[i / SCAN_TILE]; out[i] = v; if (cursors) cursors[i] = v; }
  if (i == 0) out[n] = tile_sums[ntiles];
}

// per-window maximum bucket population (decides the number of tree rounds); out[w] must be zeroed
B200_KERNEL void k_window_max(const uint32_t* __restrict__ counts, uint32_t B, uint32_t* __restrict__ out) {
  uint32_t w = blockIdx.y;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = (i < B) ? counts[(uint64_t)w * B + i] : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v) atomicMax(out + w, v);
}

// ------------------------------------------------------------------ bucket reduction by in-place folding
// Level with block size s: for every aligned block [k*s, (k+1)*s) of every window: lower half += upper half.
// After all logB levels:  sum_{b=1..B} b*T[b-1] = T[0] + sum_{j<logB} 2^j * T[2^j]   (see DESIGN.md).
// Only blocks {0, 1, 2, 4, ...} are live at each level; the others are never read again and are skipped.
template <class C>
__global__ void __launch_bounds__(128) k_fold(void* __restrict__ buckets, uint32_t W, uint32_t B, uint32_t s) {
  uint32_t half = s >> 1;
  uint32_t nblk = B / s;                       // blocks per window at this level
  uint32_t live = 1 + (nblk > 1 ? 32 - __clz(nblk - 1) : 0);   // {0} U {2^k < nblk}
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t per_win = (uint64_t)live * half;
  if (t >= per_win * W) return;
  uint32_t w = (uint32_t)(t / per_win);
  uint32_t r = (uint32_t)(t % per_win);
  uint32_t lb = r / half, i = r % half;
  uint32_t blk = lb == 0 ? 0 : (1u << (lb - 1));
  uint64_t lo = (uint64_t)w * B + (uint64_t)blk * s + i;
  XYZZ<C> a, b;
  xyzz_load<C>(a, buckets, lo); xyzz_load<C>(b, buckets, lo + half);
  xyzz_add<C>(a, b);
  xyzz_store<C>(buckets, lo, a);
}

// The last levels of the folding in ONE launch.  After the levels with block size > TAIL have run, every live block of
// TAIL buckets is an independent instance of the same problem (everything that happens to an aligned block stays inside it),
// so one CTA takes one live block and runs its remaining log2(TAIL) levels with a CTA barrier between levels instead of a
// kernel launch (these levels are latency-bound: one XYZZ addition deep each).
constexpr uint32_t FOLD_TAIL = 1024;
template <class C>
__global__ void __launch_bounds__(256) k_fold_tail(void* __restrict__ buckets, uint32_t B, uint32_t tail, uint32_t live_blocks) {
  const uint32_t w = blockIdx.x / live_blocks, lb0 = blockIdx.x % live_blocks;
  const uint32_t blk0 = lb0 == 0 ? 0 : (1u << (lb0 - 1));
  const uint64_t base = (uint64_t)w * B + (uint64_t)blk0 * tail;
  uint32_t live = 1;
  for (uint32_t sz = tail; sz >= 2; sz >>= 1, live++) {
    const uint32_t half = sz >> 1, work = live * half;
    for (uint32_t r = threadIdx.x; r < work; r += blockDim.x) {
      const uint32_t lb = r / half, i = r % half;
      const uint32_t blk = lb == 0 ? 0 : (1u << (lb - 1));
      const uint64_t lo = base + (uint64_t)blk * sz + i;
      XYZZ<C> a, b;
      xyzz_load<C>(a, buckets, lo); xyzz_load<C>(b, buckets, lo + half);
      xyzz_add<C>(a, b);
      xyzz_store<C>(buckets, lo, a);
    }
    __syncthreads();
  }
}

// The same levels spread over a thread-block CLUSTER (8 CTAs of 64 threads on 8 SMs, one live block of `tail` buckets per cluster).
// One CTA of 256 threads puts 2 warps on every scheduler of ONE SM and the rest of the GPU idles: a level is one XYZZ addition deep
// (14 multiplications = 4200 dependent IMAD.WIDE), i.e. bound by that SM's multiplier pipe at 2 warps per scheduler, and the first three
// levels hold two additions per thread (measured: 235 us for the ten levels of a 1024-bucket block, of which the last groups' are the
// unoverlapped end of every MSM).  Across a cluster every level is one addition per thread at less than one warp per scheduler; the
// barrier between levels is the hardware cluster barrier, operands are read through L2 (another SM wrote them).
constexpr uint32_t FOLD_CLUSTER = 8, FOLD_CLUSTER_THREADS = 64;
template <class C> B200_DI void xyzz_load_l2(XYZZ<C>& p, const void* base, uint64_t idx) {
  const char* s = reinterpret_cast<const char*>(base) + idx * (uint64_t)(16 * C::N);
  fe_load_l2<C>(p.x, s); fe_load_l2<C>(p.y, s + 4 * C::N); fe_load_l2<C>(p.zz, s + 8 * C::N); fe_load_l2<C>(p.zzz, s + 12 * C::N);
}
B200_DI uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
B200_DI void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <class C>
__global__ void __launch_bounds__(FOLD_CLUSTER_THREADS) k_fold_tail_cluster(void* __restrict__ buckets, uint32_t B, uint32_t tail, uint32_t live_blocks) {
  const uint32_t cl = blockIdx.x / FOLD_CLUSTER;                      // cluster = FOLD_CLUSTER consecutive CTAs of the 1-D grid
  const uint32_t w = cl / live_blocks, lb0 = cl % live_blocks;
  const uint32_t blk0 = lb0 == 0 ? 0 : (1u << (lb0 - 1));
  const uint64_t base = (uint64_t)w * B + (uint64_t)blk0 * tail;
  const uint32_t tid = cluster_ctarank() * FOLD_CLUSTER_THREADS + threadIdx.x, nthr = FOLD_CLUSTER * FOLD_CLUSTER_THREADS;
  uint32_t live = 1;
  for (uint32_t sz = tail; sz >= 2; sz >>= 1, live++) {
    const uint32_t half = sz >> 1, work = live * half;
    for (uint32_t r = tid; r < work; r += nthr) {
      const uint32_t lb = r / half, i = r % half;
      const uint32_t blk = lb == 0 ? 0 : (1u << (lb - 1));
      const uint64_t lo = base + (uint64_t)blk * sz + i;
      XYZZ<C> a, b;
      xyzz_load_l2<C>(a, buckets, lo); xyzz_load_l2<C>(b, buckets, lo + half);
      xyzz_add<C>(a, b);
      xyzz_store<C>(buckets, lo, a);
    }
    __threadfence();
    cluster_barrier();
  }
}

// ---- the same levels with FOUR LANES PER ADDITION (Fq curves) ---------------------------------------------------------------------------------
// MEASURED AND NOT ADOPTED (round 2, profiles/README.md r2late; -DB200_EXPERIMENTS builds, option fold_cluster = 2): bit-exact, but the tail kernel takes 212 us
// instead of 229 at 2^16 (CUPTI, four lanes) and the MSM is unchanged or slower (2^16 1.447 vs 1.451, 2^18 2.58 vs 2.56, 2^20 6.32 vs 6.24 ms): a level is bound by
// the L2 round trips and the cluster barrier between levels at least as much as by the addition's multiplier chain.
#if defined(B200_EXPERIMENTS)
// A level of the tail is one XYZZ addition deep, and one thread's addition is a chain of 14 dependent multiplications (~1.1 us each: the
// carry chains issue in order) = 16 us per level, 160 us for the ten levels -- the unoverlapped end of every MSM.  The addition's data flow
// is only FOUR multiplications deep:
//   depth 1: U1 = X1*ZZ2   U2 = X2*ZZ1   S1 = Y1*ZZZ2   S2 = Y2*ZZZ1          P = U2 - U1, R = S2 - S1
//   depth 2: PP = P*P      RR = R*R      ZZ12 = ZZ1*ZZ2  ZZZ12 = ZZZ1*ZZZ2
//   depth 3: Q = U1*PP     PPP = P*PP    ZZ3 = ZZ12*PP                         X3 = RR - PPP - 2Q
//   depth 4: t = R*(Q-X3)  SP = S1*PPP   ZZZ3 = ZZZ12*PPP                      Y3 = t - SP
// (g1m_add, build_curve_jacobian_a0.js:541-658, in the XYZZ form of xyzz_add).  The four lanes of a quad hold the same two points, multiply
// the four operand pairs of a depth at the same time (operands picked by lane role with selects: no divergence) and exchange the products by
// shuffles inside the quad (12 words per value, ~50 shuffles per depth against ~1600 multiplier instructions).  Same field values as the
// one-thread addition, bit for bit.  The special cases are decided identically by all four lanes (they hold the same data).
template <class C> B200_DI void quad_gather(Fe<C::N> (&g)[4], const Fe<C::N>& mine, uint32_t qmask, uint32_t qbase) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int i = 0; i < C::N; i++) g[k].l[i] = __shfl_sync(qmask, mine.l[i], qbase + k);
  }
}
template <class C> B200_DI void fe_sel4(Fe<C::N>& r, uint32_t role, const Fe<C::N>& a0, const Fe<C::N>& a1, const Fe<C::N>& a2, const Fe<C::N>& a3) {
#pragma unroll
  for (int i = 0; i < C::N; i++) r.l[i] = role == 0 ? a0.l[i] : role == 1 ? a1.l[i] : role == 2 ? a2.l[i] : a3.l[i];
}
template <class C> B200_DI void xyzz_add_quad(XYZZ<C>& acc, const XYZZ<C>& q, uint32_t role, uint32_t qmask, uint32_t qbase) {
  if (xyzz_is_inf<C>(q)) return;
  if (xyzz_is_inf<C>(acc)) { acc = q; return; }
  Fe<C::N> a, b, m, g[4];
  fe_sel4<C>(a, role, acc.x, q.x, acc.y, q.y); fe_sel4<C>(b, role, q.zz, acc.zz, q.zzz, acc.zzz);
  fe_mul<C>(m, a, b); quad_gather<C>(g, m, qmask, qbase);
  Fe<C::N> U1 = g[0], S1 = g[2], P, R;
  fe_sub<C>(P, g[1], g[0]); fe_sub<C>(R, g[3], g[2]);
  if (fe_is_zero<C>(P)) {
    if (fe_is_zero<C>(R)) { XYZZ<C> d; xyzz_dbl<C>(d, q); acc = d; } else xyzz_set_inf<C>(acc);
    return;
  }
  fe_sel4<C>(a, role, P, R, acc.zz, acc.zzz); fe_sel4<C>(b, role, P, R, q.zz, q.zzz);
  fe_mul<C>(m, a, b); quad_gather<C>(g, m, qmask, qbase);
  Fe<C::N> PP = g[0], RR = g[1], ZZ12 = g[2], ZZZ12 = g[3];
  fe_sel4<C>(a, role, U1, P, ZZ12, U1);
  fe_mul<C>(m, a, PP); quad_gather<C>(g, m, qmask, qbase);
  Fe<C::N> Q = g[0], PPP = g[1], X3, t;
  acc.zz = g[2];
  fe_sub<C>(X3, RR, PPP); fe_sub<C>(X3, X3, Q); fe_sub<C>(X3, X3, Q);
  fe_sub<C>(t, Q, X3);
  fe_sel4<C>(a, role, R, S1, ZZZ12, ZZZ12); fe_sel4<C>(b, role, t, PPP, PPP, PPP);
  fe_mul<C>(m, a, b); quad_gather<C>(g, m, qmask, qbase);
  acc.x = X3; fe_sub<C>(acc.y, g[0], g[1]); acc.zzz = g[2];
}
constexpr uint32_t FOLD_QUAD_THREADS = 128;      // 8 CTAs x 128 threads = 256 quads per live block
template <class C>
__global__ void __launch_bounds__(FOLD_QUAD_THREADS) k_fold_tail_quad(void* __restrict__ buckets, uint32_t B, uint32_t tail, uint32_t live_blocks) {
  const uint32_t cl = blockIdx.x / FOLD_CLUSTER;
  const uint32_t w = cl / live_blocks, lb0 = cl % live_blocks;
  const uint32_t blk0 = lb0 == 0 ? 0 : (1u << (lb0 - 1));
  const uint64_t base = (uint64_t)w * B + (uint64_t)blk0 * tail;
  const uint32_t tid = cluster_ctarank() * FOLD_QUAD_THREADS + threadIdx.x, nquads = FOLD_CLUSTER * FOLD_QUAD_THREADS / 4;
  const uint32_t quad = tid >> 2, role = tid & 3, lane = threadIdx.x & 31, qbase = lane & ~3u, qmask = 0xfu << qbase;
  uint32_t live = 1;
  for (uint32_t sz = tail; sz >= 2; sz >>= 1, live++) {
    const uint32_t half = sz >> 1, work = live * half;
    for (uint32_t r = quad; r < work; r += nquads) {
      const uint32_t lb = r / half, i = r % half;
      const uint32_t blk = lb == 0 ? 0 : (1u << (lb - 1));
      const uint64_t lo = base + (uint64_t)blk * sz + i;
      XYZZ<C> a, b;
      xyzz_load_l2<C>(a, buckets, lo); xyzz_load_l2<C>(b, buckets, lo + half);
      xyzz_add_quad<C>(a, b, role, qmask, qbase);
      Fe<C::N> mine; fe_sel4<C>(mine, role, a.x, a.y, a.zz, a.zzz);      // lane r stores coordinate r
      fe_store<C>(reinterpret_cast<char*>(buckets) + lo * (uint64_t)(16 * C::N) + (uint64_t)role * (4 * C::N), mine);
    }
    __threadfence();
    cluster_barrier();
  }
}
#endif  // B200_EXPERIMENTS

// One thread per slot: R_w = T[0] + sum_j 2^j T[2^j] by Horner over j (logB doublings).  The extra slot (index Wd,
// present when W == Wd + 1) holds buckets B+1 .. 2B of the last window: its value is the same expression + B * T[0].
template <class C>
__global__ void __launch_bounds__(32) k_window_sums(const void* __restrict__ buckets, uint32_t W, uint32_t Wd, uint32_t B, uint32_t logB, void* __restrict__ wsum) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  XYZZ<C> acc; xyzz_set_inf<C>(acc);
  if (w >= Wd) xyzz_load<C>(acc, buckets, (uint64_t)w * B);          // + B*T[0] = 2^logB * T[0]: seed the Horner chain one step higher
  for (int j = (int)logB - 1; j >= 0; j--) {
    if (w >= Wd || j + 1 < (int)logB) { XYZZ<C> d; xyzz_dbl<C>(d, acc); acc = d; }
    XYZZ<C> t; xyzz_load<C>(t, buckets, (uint64_t)w * B + (1u << j));
    xyzz_add<C>(acc, t);
  }
  XYZZ<C> t0; xyzz_load<C>(t0, buckets, (uint64_t)w * B);
  xyzz_add<C>(acc, t0);
  xyzz_store<C>(wsum, w, acc);
}

// result = sum_w 2^(c*w) R_w, top window first (accumulateAcrossChunks, build_multiexp_opt.js:1710-1746;
// multiexp Horner loop, build_multiexp.js:319-369).  Output: Jacobian Montgomery x||y||z, canonical zero for infinity.
// GPU form of the window combination; a single dependent chain of Wd*c doublings (see DESIGN.md for why the engine's
// default performs this last serial step on the host from the folded bucket arrays instead).
template <class C>
__global__ void __launch_bounds__(32) k_horner(const void* __restrict__ wsum, uint32_t W, uint32_t Wd, uint32_t c0, uint32_t rem, void* __restrict__ out_jac) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  XYZZ<C> acc; xyzz_set_inf<C>(acc);
  for (int w = (int)Wd - 1; w >= 0; w--) {
    const uint32_t cw = c0 + ((uint32_t)w < rem ? 1u : 0u);       // acc holds windows > w: shift it by the width of window w
    if (!xyzz_is_inf<C>(acc)) for (uint32_t k = 0; k < cw; k++) { XYZZ<C> d; xyzz_dbl<C>(d, acc); acc = d; }
    XYZZ<C> t; xyzz_load<C>(t, wsum, w);
    xyzz_add<C>(acc, t);
    if (w + 1 == (int)Wd && W > Wd) { xyzz_load<C>(t, wsum, Wd); xyzz_add<C>(acc, t); }
  }
  Fe<C::N> X, Y, Z;
  xyzz_to_jacobian<C>(X, Y, Z, acc);
  char* o = reinterpret_cast<char*>(out_jac);
  fe_store<C>(o, X); fe_store<C>(o + 4 * C::N, Y); fe_store<C>(o + 8 * C::N, Z);
}

// Collect the logB + 1 live entries of every folded slot (T[0], T[1], T[2], T[4], ...) into a compact array for the
// host-side window combination: out[w * (logB + 1) + 0] = T_w[0], out[w * (logB + 1) + 1 + j] = T_w[2^j].
template <class C>
__global__ void k_gather_folded(const void* __restrict__ buckets, uint32_t W, uint32_t B, uint32_t logB, void* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t per = logB + 1;
  if (t >= W * per) return;
  uint32_t w = t / per, k = t % per;
  uint64_t src = (uint64_t)w * B + (k == 0 ? 0u : (1u << (k - 1)));
  XYZZ<C> p; xyzz_load<C>(p, buckets, src);
  xyzz_store<C>(out, t, p);
}

// ------------------------------------------------------------------ precomputed window tables (resident bases only)
// table row w holds 2^(bit offset of window w) * P_i for every resident base, so the digits of ALL windows can share ONE
// bucket array: the per-window bucket reductions and the window combination (W*c doublings) disappear from the MSM, and
// the window can be wider than a per-window bucket array could afford.  Built once at upload time: the running multiples
// are kept in XYZZ and doubled in place (g1m_double, build_curve_jacobian_a0.js:291-359), each row is converted to affine
// with k_xyzz_to_affine.  The reference has no such table (it receives the bases with every call); the sum is unchanged.
template <class C>
__global__ void __launch_bounds__(128) k_table_init(const void* __restrict__ bases, uint32_t n, void* __restrict__ cur_xyzz) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> p; affine_load<C>(p, bases, i);
  XYZZ<C> q;
  if (affine_is_inf<C>(p)) xyzz_set_inf<C>(q); else xyzz_from_affine<C>(q, p);
  xyzz_store<C>(cur_xyzz, i, q);
}
template <class C>
__global__ void __launch_bounds__(128) k_table_double(void* __restrict__ cur_xyzz, uint32_t n, uint32_t doublings) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<C> p; xyzz_load<C>(p, cur_xyzz, i);
  for (uint32_t k = 0; k < doublings; k++) { XYZZ<C> d; xyzz_dbl<C>(d, p); p = d; }
  xyzz_store<C>(cur_xyzz, i, p);
}

// ------------------------------------------------------------------ small utility kernels behind the C ABI
// g1m_normalize + f1m_fromMontgomery x2 (build_curve_jacobian_a0.js:940-973; test/batchAffine.js:1249-1254):
// Jacobian Montgomery -> canonical affine x||y as plain integers; infinity -> zeros.
template <class C>
__global__ void k_normalize(const void* __restrict__ jac, void* __restrict__ xy, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const char* s = reinterpret_cast<const char*>(jac) + (uint64_t)i * 12 * C::N;
  char* d = reinterpret_cast<char*>(xy) + (uint64_t)i * 8 * C::N;
  Fe<C::N> X, Y, Z, zi, zi2, zi3;
  fe_load<C>(X, s); fe_load<C>(Y, s + 4 * C::N); fe_load<C>(Z, s + 8 * C::N);
  if (fe_is_zero<C>(Z)) { fe_set_zero<C>(X); fe_store<C>(d, X); fe_store<C>(d + 4 * C::N, X); return; }
  fe_inv_fast<C>(zi, Z); fe_sqr<C>(zi2, zi); fe_mul<C>(zi3, zi2, zi);
  fe_mul<C>(X, X, zi2); fe_mul<C>(Y, Y, zi3);
  fe_from_mont<C>(X, X); fe_from_mont<C>(Y, Y);
  fe_store<C>(d, X); fe_store<C>(d + 4 * C::N, Y);
}

// out = sum of n Jacobian points (g1m_add) -- combines per-GPU partial results.  One warp: lane t adds up points t, t + 32, ...,
// then a shared-memory tree over the lanes; the dependent chain is ceil(n / 32) + log2(min(n, 32)) additions instead of n (8 partial
// results: 3 deep instead of 8 -- the chain is pure latency, ~17 us per XYZZ addition on one thread, and sits inside every multi-GPU step).
template <class C>
__global__ void __launch_bounds__(32) k_sum_jacobian(const void* __restrict__ pts, uint32_t n, void* __restrict__ out_jac) {
  __shared__ uint32_t sh[16 * 16 * C::N];      // lanes w .. 2w-1 hand their partial sums to lanes 0 .. w-1 (slot = lane - w)
  const uint32_t t = threadIdx.x;
  XYZZ<C> acc; xyzz_set_inf<C>(acc);
  for (uint32_t i = t; i < n; i += 32) {
    const char* s = reinterpret_cast<const char*>(pts) + (uint64_t)i * 12 * C::N;
    Fe<C::N> X, Y, Z; fe_load_cg<C>(X, s); fe_load_cg<C>(Y, s + 4 * C::N); fe_load_cg<C>(Z, s + 8 * C::N);
    XYZZ<C> p;
    if (fe_is_zero<C>(Z)) xyzz_set_inf<C>(p);
    else { p.x = X; p.y = Y; fe_sqr<C>(p.zz, Z); fe_mul<C>(p.zzz, p.zz, Z); }
    xyzz_add<C>(acc, p);
  }
  for (uint32_t w = 16; w >= 1; w >>= 1) {
    if (t >= w && t < 2 * w) {
#pragma unroll
      for (int k = 0; k < C::N; k++) { sh[((t - w) * 4 + 0) * C::N + k] = acc.x.l[k]; sh[((t - w) * 4 + 1) * C::N + k] = acc.y.l[k]; sh[((t - w) * 4 + 2) * C::N + k] = acc.zz.l[k]; sh[((t - w) * 4 + 3) * C::N + k] = acc.zzz.l[k]; }
    }
    __syncwarp();
    if (t < w && t + w < n) {
      XYZZ<C> q;
#pragma unroll
      for (int k = 0; k < C::N; k++) { q.x.l[k] = sh[(t * 4 + 0) * C::N + k]; q.y.l[k] = sh[(t * 4 + 1) * C::N + k]; q.zz.l[k] = sh[(t * 4 + 2) * C::N + k]; q.zzz.l[k] = sh[(t * 4 + 3) * C::N + k]; }
      xyzz_add<C>(acc, q);
    }
    __syncwarp();
  }
  if (t != 0) return;
  Fe<C::N> X, Y, Z; xyzz_to_jacobian<C>(X, Y, Z, acc);
  char* o = reinterpret_cast<char*>(out_jac);
  fe_store<C>(o, X); fe_store<C>(o + 4 * C::N, Y); fe_store<C>(o + 8 * C::N, Z);
}

// Synthetic bases: P_i = k_i * G with k_i = splitmix64(seed + first + i) -- the same stream as
// oracle_point_scalar() in oracle/msm_oracle.c so that CPU and GPU generate identical inputs.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <class C>
__global__ void __launch_bounds__(128) k_generate_xyzz(const void* __restrict__ gen_xy, uint64_t seed, uint64_t first, uint32_t n, void* __restrict__ out_xyzz) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = splitmix64(seed + first + i); if (!k) k = 1;
  Affine<C> g; affine_load<C>(g, gen_xy, 0);
  XYZZ<C> acc; xyzz_set_inf<C>(acc);
  for (int b = 63; b >= 0; b--) {
    XYZZ<C> d; xyzz_dbl<C>(d, acc); acc = d;
    if ((k >> b) & 1) xyzz_madd<C>(acc, g);
  }
  xyzz_store<C>(out_xyzz, i, acc);
}
// XYZZ -> affine, one inversion per thread amortised over GROUP points (Montgomery trick, build_batchinverse.js:4-140)
template <class C, int GROUP>
__global__ void __launch_bounds__(128) k_xyzz_to_affine(const void* __restrict__ in_xyzz, uint32_t n, void* __restrict__ out_xy) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t base = (uint64_t)t * GROUP;
  if (base >= n) return;
  Fe<C::N> pre[GROUP], acc; fe_set_one<C>(acc);
  uint32_t cnt = (n - base < GROUP) ? (uint32_t)(n - base) : GROUP;
  for (uint32_t k = 0; k < cnt; k++) {
    XYZZ<C> p; xyzz_load<C>(p, in_xyzz, base + k);
    pre[k] = acc;
    if (!xyzz_is_inf<C>(p)) fe_mul<C>(acc, acc, p.zzz);
  }
  Fe<C::N> inv; fe_inv<C>(inv, acc);
  for (int k = (int)cnt - 1; k >= 0; k--) {
    XYZZ<C> p; xyzz_load<C>(p, in_xyzz, base + k);
    Affine<C> a;
    if (xyzz_is_inf<C>(p)) { affine_set_inf<C>(a); }
    else {
      Fe<C::N> zi3, zi2; fe_mul<C>(zi3, inv, pre[k]); fe_mul<C>(inv, inv, p.zzz);
      // 1/zz = zzz^-1 * ... : zz^3 = zzz^2  =>  1/zz = (zz / zzz)^2 ... use 1/zz = zi3^2 * zz^2 (since zi3 = 1/zzz, zz^3 = zzz^2)
      fe_mul<C>(zi2, zi3, p.zz); fe_sqr<C>(zi2, zi2);
      fe_mul<C>(a.x, p.x, zi2); fe_mul<C>(a.y, p.y, zi3);
    }
    affine_store<C>(out_xy, base + k, a);
  }
}

// elementwise field ops for kernel-level parity tests (op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv, 5 toMont, 6 fromMont, 7 neg, 8 inv by Fermat)
template <class C>
__global__ void k_fp_op(int op, const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ r, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<C::N> x, y, z;
  fe_load<C>(x, reinterpret_cast<const char*>(a) + (uint64_t)i * 4 * C::N);
  if (b) fe_load<C>(y, reinterpret_cast<const char*>(b) + (uint64_t)i * 4 * C::N); else fe_set_zero<C>(y);
  switch (op) {
    case 0: fe_mul<C>(z, x, y); break;
    case 1: fe_add<C>(z, x, y); break;
    case 2: fe_sub<C>(z, x, y); break;
    case 3: fe_sqr<C>(z, x); break;
    case 4: fe_inv_fast<C>(z, x); break;
    case 5: fe_to_mont<C>(z, x); break;
    case 6: fe_from_mont<C>(z, x); break;
    case 7: fe_neg<C>(z, x); break;
    default: fe_inv<C>(z, x); break;      // 8: Fermat inversion (cross-check of the binary-Euclid one)
  }
  fe_store<C>(reinterpret_cast<char*>(r) + (uint64_t)i * 4 * C::N, z);
}

// Register-resident integer-multiply throughput probes: the roofline denominators for the accumulate phase
// ("Peak = measured on the box by a register-resident mad.wide microbenchmark").
// What was learnt writing them (all verified in SASS with cuobjdump):
//  * a loop-invariant product is hoisted and the "multiply-adds" become IADD3s -- every multiply below takes one
//    operand from its own accumulator;
//  * mad.wide.u32 without a carry is NOT kept as one instruction: ptxas splits it into IMAD.WIDE.U32 (RZ addend) +
//    IADD3/IADD3.X, and a mul.wide whose high half is unused becomes a 32-bit IMAD;
//  * the only form that stays a single 32x32+64 instruction is the carry-chain one (mad.lo.cc/madc.hi.cc ->
//    IMAD.WIDE.U32[.X] with predicate carry), which is exactly what the Montgomery multiplier in fp.cuh is made of.
// Measured on B200: 32-bit IMAD 18.5e12 /s (64 lanes/clk/SM); IMAD.WIDE.U32 carry chains 9.18e12 /s (32 lanes/clk/SM).
// The second number is the peak of the instruction the field arithmetic uses: the roofline denominator.
template <int MODE>   // 1: IMAD (32-bit, acc = acc * b + a), 8 independent chains
__global__ void __launch_bounds__(256) k_imad_probe(uint32_t iters, uint32_t seed, unsigned long long* __restrict__ sink) {
  uint32_t a = seed + threadIdx.x, b = (seed * 3 + blockIdx.x) | 1u;
  uint32_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = a * (2 * k + 3) + b;
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(b), "r"(a));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  if (s == 0x1234567u) sink[0] = s;
}
// IMAD.WIDE.U32 with carry in/out: 4 independent carry chains of 4 wide multiply-adds each (16 per inner iteration).
B200_KERNEL void __launch_bounds__(256) k_imadx_probe(uint32_t iters, uint32_t seed, uint32_t* __restrict__ sink) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t r[32];
#pragma unroll
  for (int k = 0; k < 32; k++) r[k] = a + k;
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t* x = r + 8 * c;
      mad_lo_cc(x[0], a, b, x[0]); madc_hi_cc(x[1], a, b, x[1]);
      madc_lo_cc(x[2], a, b, x[2]); madc_hi_cc(x[3], a, b, x[3]);
      madc_lo_cc(x[4], a, b, x[4]); madc_hi_cc(x[5], a, b, x[5]);
      madc_lo_cc(x[6], a, b, x[6]); madc_hi(x[7], a, b, x[7]);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 32; k++) s ^= r[k];
  if (s == 0x1234567u) sink[0] = s;
}
// Field-multiply throughput probe: ITER dependent Montgomery multiplications per thread (2N^2+N limb products each).
#if defined(B200_EXPERIMENTS)
// FP64 pipe probe: 8 independent DFMA chains per thread (the pipe an FP64-based multiplier would run on, beside the integer pipe)
B200_KERNEL void __launch_bounds__(256) k_dfma_probe(uint32_t iters, double seed, double* __restrict__ sink) {
  double a[8], b = seed * 1.0000001, c = seed * 0.5;
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed + k;
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = __fma_rz(a[k], b, c);
  }
  double t = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) t += a[k];
  if (t == 12345.678) sink[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
#endif
template <class C>
__global__ void __launch_bounds__(256) k_fpmul_probe(uint32_t iters, const void* __restrict__ in, void* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe<C::N> x, y;
  fe_load<C>(x, reinterpret_cast<const char*>(in) + (uint64_t)(i & 1023) * 4 * C::N);
  y = x;
  if (iters & 0x80000000u) { for (uint32_t k = 0; k < (iters & 0x7fffffffu); k++) { fe_sqr<C>(y, y); } }
  else for (uint32_t k = 0; k < iters; k++) { fe_mul<C>(y, y, x); }
  fe_store<C>(reinterpret_cast<char*>(out) + (uint64_t)i * 4 * C::N, y);
}

}  // namespace b200
