// b200msm.cu -- context, pipeline orchestration and the C ABI declared in include/b200msm.h.
//
// Host-side counterpart of the orchestration functions of the reference:
//   g1m_multiexpAffine       wasmcurves/src/build_multiexp.js:251-371
//   g1m_multiexpAffine_chunk wasmcurves/src/build_multiexp.js:96-249
//   g1m_multiexp_multiExp    wasmcurves/src/build_multiexp_opt.js:1987-2110 (scratch allocation :2037-2070)
// The reference bump-allocates its scratch in WASM linear memory per call; here the context owns
// grow-only device buffers that are reused across calls (no cudaMalloc in steady state).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <map>
#include <vector>
#include <algorithm>

#include "../../include/b200msm.h"
#include "../../include/b200msm_probes.h"
#include "msm_kernels.cuh"
#include "accumulate.cuh"
#if defined(B200_EXPERIMENTS)
#include "fp29.cuh"
#endif
#include "codecs.cuh"
#include "glv.cuh"
#include "host_ec.h"
#include "internal.h"
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <functional>
#include <atomic>
#include <memory>

using namespace b200;

namespace {

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    // the new block is obtained BEFORE the old one is released: a failed allocation leaves the buffer as it was.  Only when the two
    // do not fit side by side is the old block given up first (grow-only scratch holds no data across calls).
    size_t want = bytes + (bytes >> 3) + 256;
    void* np_ = nullptr;
    cudaError_t e = cudaMalloc(&np_, want);
    if (e != cudaSuccess && p) {
      cudaGetLastError(); cudaFree(p); p = nullptr; cap = 0;
      e = cudaMalloc(&np_, want);
    }
    if (e != cudaSuccess) return e;
    if (p) cudaFree(p);
    p = np_; cap = want; return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Resident { int curve; uint64_t n; void* d; void* dx = nullptr;      // dx: the x coordinates alone (built at upload for sets of >= 2^19 points, see k_extract_x)
                                 // d: n affine points; with a window table: Wd rows of n points, row 0 = the bases
                  uint32_t t_nbits = 0, t_c0 = 0, t_rem = 0, t_Wd = 0; };   // window plan the table was built for (t_Wd == 0: no table)
struct Precomp { uint32_t stride, nbits, c0, rem, Wd; };
// one base set spread over the devices of a multi context: device g holds points [lo[g], lo[g] + cnt[g]) under its own handle h[g]
// (sharded: disjoint ranges; replicated: every device holds [0, n))
struct MultiResident { int curve; uint64_t n; bool replicated; std::vector<uint64_t> h, lo, cnt; };

// one accumulate lane: a stream with its own tree scratch (see accumulate_batch_affine)
struct TreeLane {
  cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;
  DevBuf offs, tiles, bid, pa, pb, prefix, prod, lvlprefix, others, meta, carry, btree;
};
constexpr int MAX_LANES = 8;

// ---- issue threads: one host thread per lane -----------------------------------------------------------------------------------------
// A window group is ~40-200 kernel launches (rounds x {operand table, forward pass, product-tree levels, root, backward pass}, fold levels).
// Issued from ONE host thread the lanes start one after the other, each ~0.1 ms (its launches) behind the previous one; below ~2^19 points
// that stagger would be a tenth of the whole MSM.  With one issuing thread per lane all lanes start together.  MEASURED (profiles/README.md r2): the
// stagger seen in a CUPTI trace is the tracer's own per-launch cost; untraced, a launch costs ~2 us and the option changes nothing
// (2^16: 1.53 / 1.58 ms, 2^18: 2.63 / 2.64, 2^20: 6.57 / 6.56 with 0 / 1) -- kept as the option "issue_threads", off by default.
// The threads only enqueue work on their lane's stream; everything shared is read-only while they run.
struct IssuePool {
  struct Worker { std::thread th; std::mutex m; std::condition_variable cv; std::deque<std::function<void()>> q; std::atomic<int> pending{0}; bool quit = false; };
  std::vector<std::unique_ptr<Worker>> w;
  void ensure(int n, int device) {
    while ((int)w.size() < n) {
      w.emplace_back(new Worker()); Worker* W = w.back().get();
      W->th = std::thread([W, device] {
        cudaSetDevice(device);
        for (;;) {
          for (int spin = 0; spin < 4000 && W->pending.load(std::memory_order_acquire) == 0; spin++) { }      // a job usually follows within microseconds of the previous one
          std::function<void()> f;
          { std::unique_lock<std::mutex> lk(W->m);
            W->cv.wait(lk, [&] { return W->quit || !W->q.empty(); });
            if (W->q.empty()) return;
            f = std::move(W->q.front()); W->q.pop_front(); W->pending.fetch_sub(1, std::memory_order_relaxed); }
          f();
        }
      });
    }
  }
  void submit(int lane, std::function<void()> f) {
    Worker* W = w[lane].get();
    { std::lock_guard<std::mutex> lk(W->m); W->q.push_back(std::move(f)); W->pending.fetch_add(1, std::memory_order_release); }
    W->cv.notify_one();
  }
  ~IssuePool() {
    for (auto& p : w) { { std::lock_guard<std::mutex> lk(p->m); p->quit = true; } p->cv.notify_one(); }
    for (auto& p : w) if (p->th.joinable()) p->th.join();
  }
};
constexpr uint64_t WARP_LEVEL_MAX = 131072;     // product-tree levels with at most this many values use the warp-assisted kernel (only one such level can occur: 131072 / 128 <= BA_ROOT_MAX)

}  // namespace

struct b200msm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr; bool own_stream = false;
  std::string err;
  int opt_window_bits = 0, opt_accumulate = 0, opt_tree_rounds = -1;
  DevBuf bases, scalars, canon, counts, offsets, ranks, tiles, sorted, buckets, wsum, out, misc, acc_a, acc_b, acc_c, acc_d, acc_e, jac_in, jac_affine;
  TreeLane lane[MAX_LANES];                                                   // batch-affine tree lanes
  int opt_sort_groups = 1;                                // sort the window slots group by group on the lanes' streams (run_grouped)
  int opt_lanes = 4, opt_ba_k = 0, opt_pt_k = 8, opt_persist = 592, opt_subslots = 0, opt_probe_smem = 0;
  bool probe29 = false, probe_sqr = false; int64_t opt_group_pairs = 0;
  cudaEvent_t ev_plan = nullptr, ev_sorted = nullptr, ev_bases = nullptr, ev_done = nullptr;
  cudaStream_t copy_stream = nullptr; bool bases_pending = false;
  DevBuf xonly; const void* xs_hint = nullptr; const void* cur_xs = nullptr; uint64_t cur_xs_bytes = 0; int opt_xonly = 1; bool prof_noxs = false;      // x coordinates of the bases alone (k_extract_x) for the forward pass of round 0
  std::vector<cudaEvent_t> gev;                                               // one event per window group (folded points on the host)
  uint32_t* h_pinned = nullptr;
  void* h_folded = nullptr; size_t h_folded_cap = 0;                         // pinned staging of the folded bucket entries
  int opt_combine = 0;                                                        // 0 = host serial tail (default), 1 = device k_window_sums + k_horner
  float host_combine_ms = 0;                                               // small pinned read-back area (2048 words: [0,512) slot maxima, [512,1024) slot pair offsets, [1024,..) per-round totals)
  std::map<uint64_t, Resident> residents; uint64_t next_handle = 1;
  cudaEvent_t ev[8] = {};
  // fine-grained phase profiler (active only while a stats struct is being filled)
  std::vector<cudaEvent_t> pev; std::vector<int> ptag; size_t pused = 0; bool prof = false;
  std::atomic<uint64_t> launches{0}; uint64_t adds_r0 = 0, adds_exact = 0, cur_n = 0;
  IssuePool* pool = nullptr; int opt_issue_threads = 0, opt_fold_cluster = 1, opt_groups = 0, opt_group_small = 70, opt_persist_fwd = 0, opt_ba_k0 = 0, opt_fused = 0, opt_fused_grid = 444, opt_fused_tiles = 592, opt_fused_kmax = 16, opt_block_tree = 0, opt_meta_upfront = 0;      // measured alternatives, -DB200_EXPERIMENTS builds only (accumulate.cuh)
  int64_t opt_group_plan = 0; std::mutex err_mu;      // one issuing host thread per lane (IssuePool); err is written under err_mu
  bool blocking_waits = false;      // batch workers: host waits sleep instead of spinning (8 workers per GPU x 8 ranks would spin on more threads than the host has cores)
  size_t total_mem = 0; double mem_share = 1.0;      // fraction of the device memory budget this context may plan with (batch workers: 1 / workers)
  // multi-device context (b200msm_create_multi): devs[0] == this, devs[g] = the single-device context of device g; mres = handles of sharded / replicated base sets
  std::vector<b200msm_ctx*> devs; std::map<uint64_t, MultiResident> mres; int64_t opt_multi_min = 1 << 15; int opt_multi_replicate = 0;
  std::vector<b200msm_ctx*> workers; int opt_batch_workers = 8, opt_batch_lanes = 1, opt_batch_blocking = 0;            // sub-contexts (own stream + scratch) that run the MSMs of a batch concurrently
};

namespace {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { { std::lock_guard<std::mutex> lk_(ctx->err_mu); ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); } \
  return e_ == cudaErrorMemoryAllocation ? B200MSM_E_NOMEM : B200MSM_E_CUDA; } } while (0)
#define CKL() do { ctx->launches++; CK(cudaGetLastError()); } while (0)
enum { T_SORT = 0, T_PLAN, T_TREE_FWD, T_INV_TREE, T_TREE_BWD, T_FINISH, T_FOLD, T_WSUM, T_HORNER, T_TREE_BWD0, T_NTAGS };
#define MARK(tag) do { if (ctx->prof) { int rc_ = prof_mark(ctx, tag); if (rc_) return rc_; } } while (0)

void copy_options(b200msm_ctx* w, const b200msm_ctx* ctx) {
  w->opt_window_bits = ctx->opt_window_bits; w->opt_accumulate = ctx->opt_accumulate; w->opt_tree_rounds = ctx->opt_tree_rounds;
  w->opt_ba_k = ctx->opt_ba_k; w->opt_pt_k = ctx->opt_pt_k; w->opt_persist = ctx->opt_persist; w->opt_subslots = ctx->opt_subslots; w->opt_combine = ctx->opt_combine;
  w->opt_xonly = ctx->opt_xonly; w->opt_meta_upfront = ctx->opt_meta_upfront; w->opt_block_tree = ctx->opt_block_tree; w->opt_fused = ctx->opt_fused; w->opt_fused_grid = ctx->opt_fused_grid; w->opt_fused_tiles = ctx->opt_fused_tiles; w->opt_fused_kmax = ctx->opt_fused_kmax;
  w->opt_group_pairs = ctx->opt_group_pairs; w->opt_sort_groups = ctx->opt_sort_groups; w->opt_batch_workers = ctx->opt_batch_workers;
  // A batch worker runs whole MSMs next to other workers' MSMs: the workers are in DIFFERENT phases at any moment (one sorting, one in a
  // multiplier-bound backward pass, one in its latency-bound fold), which overlaps better than the lanes of one MSM, whose rounds run in
  // step.  So a worker uses ONE lane with full (non-persistent) grids and the plain fold tail (a cluster launch waits for 8 free SMs of one GPC).
  // Measured, 64 MSMs of 2^18 points on one B200: 4 workers x 4 lanes 133.7 ms, 8 workers x 1 lane 108.7 ms (1.70 ms per MSM); 16 x 2^20:
  // 94.0 -> 82.9 ms (5.18 ms per MSM); 64 x 2^16: 54.2 -> 44.4 ms (profiles/README.md r2).
  w->opt_lanes = ctx->opt_batch_lanes; w->opt_fold_cluster = 0; w->opt_issue_threads = 0;
}

// The groups of one MSM: job gi runs on lane gi % lanes -- inline (single issuing thread) or on that lane's issue thread.  wait(gi) blocks until
// job gi has been ISSUED (its kernels are enqueued, its completion event recorded) and returns its status; the destructor waits for every job,
// so the jobs may refer to the caller's locals.
struct GroupIssue {
  b200msm_ctx* ctx; uint32_t n; bool threaded; std::unique_ptr<std::atomic<int>[]> st; std::vector<int> rc;
  GroupIssue(b200msm_ctx* c, uint32_t ngroups, uint32_t lanes) : ctx(c), n(ngroups), st(new std::atomic<int>[ngroups]), rc(ngroups, 0) {
    for (uint32_t i = 0; i < n; i++) st[i].store(2);       // 2 = not submitted
    threaded = c->opt_issue_threads != 0 && lanes > 1 && ngroups > 1 && !c->prof;
    if (threaded) { if (!c->pool) c->pool = new IssuePool(); c->pool->ensure((int)lanes, c->device); }
  }
  template <class F> int run(uint32_t gi, uint32_t lane, F f) {
    if (!threaded) { rc[gi] = f(); st[gi].store(1); return rc[gi]; }
    st[gi].store(0);
    ctx->pool->submit((int)lane, [this, gi, f] { rc[gi] = f(); st[gi].store(1, std::memory_order_release); });
    return 0;
  }
  int wait(uint32_t gi) {
    for (int spin = 0; st[gi].load(std::memory_order_acquire) == 0; spin++) if (spin > 2000) std::this_thread::yield();
    return rc[gi];
  }
  int wait_all() { int r = 0; for (uint32_t i = 0; i < n; i++) { int x = wait(i); if (x && !r) r = x; } return r; }
  ~GroupIssue() { wait_all(); }
};

int lane_init(b200msm_ctx* ctx, TreeLane& ln) {
  if (!ln.stream) {
    int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);          // hi = numerically smallest = highest priority
    int idx = (int)(&ln - &ctx->lane[0]);
    int prio = std::min(lo, hi + idx);                                        // lane 0 highest: its kernels drain first, so the lanes' serial tails do not coincide
    CK(cudaStreamCreateWithPriority(&ln.stream, cudaStreamNonBlocking, prio));
  }
  if (!ln.done) CK(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
  return B200MSM_OK;
}

int prof_mark(b200msm_ctx* ctx, int tag) {
  if (ctx->pused == ctx->pev.size()) { cudaEvent_t e; CK(cudaEventCreate(&e)); ctx->pev.push_back(e); ctx->ptag.push_back(0); }
  ctx->ptag[ctx->pused] = tag;
  CK(cudaEventRecord(ctx->pev[ctx->pused], ctx->stream));
  ctx->pused++;
  return B200MSM_OK;
}

// device ordinal a pointer lives on, -1 for host memory (pageable, pinned or unknown)
int ptr_device(const void* p) {
  if (!p) return -1;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}
bool is_device_ptr(const void* p) { return ptr_device(p) >= 0; }

// n8 = bytes per coordinate-field element: Fq for G1 (48 / 32), Fq2 for G2 (96 / 64)
int n8_of(int curve) { static const int t[4] = {48, 32, 96, 64}; return t[curve & 3]; }
#if defined(B200_NO_G2)
bool curve_ok(int curve) { return curve >= 0 && curve <= 1; }
#else
bool curve_ok(int curve) { return curve >= 0 && curve <= 3; }
#endif
bool curve_g1(int curve) { return curve == B200MSM_BLS12_381_G1 || curve == B200MSM_BN254_G1; }
// run a statement with C bound to the field class of `curve` (G2 = the same templates over Fq2, see fp.cuh)
#if defined(B200_NO_G2)      // development builds (B200_DEV_NO_G2=1 in __graft_entry__): the two G2 instantiations are 85 % of the compile time
#define B200_CURVE_SWITCH(curve, ...) \
  switch (curve) { case 0: { using C = BLS12_381; __VA_ARGS__; } break; default: { using C = BN254; __VA_ARGS__; } break; }
#else
#define B200_CURVE_SWITCH(curve, ...) \
  switch (curve) { case 0: { using C = BLS12_381; __VA_ARGS__; } break; case 1: { using C = BN254; __VA_ARGS__; } break; \
                   case 2: { using C = Fq2<BLS12_381>; __VA_ARGS__; } break; default: { using C = Fq2<BN254>; __VA_ARGS__; } break; }
#endif

// host-side field descriptor of C for the serial tail (host_ec.h)
template <class C> struct HostField {
  static constexpr int L = C::N / 2;
  using type = b200host::Field<L>;
  static type make() {
    type f;
    for (int i = 0; i < L; i++) { f.q[i] = (uint64_t)C::q(2 * i) | ((uint64_t)C::q(2 * i + 1) << 32); f.one[i] = (uint64_t)C::one(2 * i) | ((uint64_t)C::one(2 * i + 1) << 32); }
    { uint64_t x = 1; for (int k = 0; k < 6; k++) x *= 2 - f.q[0] * x; f.np = 0 - x; }      // -q^-1 mod 2^64 (Newton)
    return f;
  }
};
template <class B> struct HostField<Fq2<B>> {
  static constexpr int L = B::N / 2;
  using type = b200host::Field2<L>;
  static type make() { type f; f.b = HostField<B>::make(); for (int i = 0; i < 2 * L; i++) f.one[i] = i < L ? f.b.one[i] : 0; return f; }
};

uint32_t auto_window_bits(uint64_t n, uint32_t nbits) {
  uint32_t lg = 0; while ((2ull << lg) <= n) lg++;          // floor(log2 n)
  int c = (int)lg - 4;
  // Small problems skip the batch-affine tree (see accumulate_batch_affine): every round costs a latency-bound inversion tail (~0.2 ms),
  // more than the whole XYZZ accumulation of 2^14 points.  Without the tree a bucket is one thread's serial chain, so the window is chosen
  // for ~4 points per bucket.  Measured (BLS12-381, tree / no tree): 2^10 1.24 / 0.74 ms, 2^12 1.19 / 0.79, 2^14 1.41 / 1.01, 2^16 1.70 / 1.55 (one round).
  if (lg <= 13) c = std::min(12, std::max(8, (int)lg + 1));
  else if (lg <= 16) c = 13;
  if (c > 16 && n < (1ull << 22)) c = 16;
  if (c > 20) c = 20;
  if (c < 2) c = 2;
  if (nbits < 2) c = 1;
  if ((uint32_t)c > nbits) c = (int)nbits;
  return (uint32_t)c;
}

// out[i] = base + sum of in[0..i), out[n] = base + total, on stream s with `tiles` as scratch
int exclusive_scan(b200msm_ctx* ctx, cudaStream_t s, DevBuf& tiles, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t base) {
  uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  CK(tiles.ensure((size_t)(ntiles + 1) * 4));
  k_scan_tiles<<<ntiles, SCAN_THREADS, 0, s>>>(in, out, n, tiles.as<uint32_t>()); CKL();
  k_scan_sums<<<1, 1024, 0, s>>>(tiles.as<uint32_t>(), ntiles, base); CKL();
  k_scan_apply<<<(n + 255) / 256, 256, 0, s>>>(out, n, tiles.as<uint32_t>(), ntiles, nullptr); CKL();
  return B200MSM_OK;
}

// vals[0 .. n) <- 1 / vals[e], in place, through the grid-wide product tree (f1m_batchInverse, build_batchinverse.js:4-140, for values that are
// all non-zero): plain K-ary levels while the level is large, one warp-assisted level (arity 32*4) once it is small, one block for the last
// <= BA_ROOT_MAX values (a single field inversion), and the same levels back down.  Levels above `vals` are appended behind it (vals must have room
// for 2*n + 4096 elements); lpre = prefix scratch of the same size; ln_.others serves the warp-assisted level.
template <class C>
int product_tree_invert(b200msm_ctx* ctx, TreeLane& ln_, cudaStream_t s, char* vals, char* lpre, uint64_t n, int PK) {
  const size_t fe = 4 * C::N;
  struct Lvl { uint64_t n; char* v; char* p; bool warp; int K; };
  std::vector<Lvl> lv;
  char* cur = vals; char* curp = lpre;
  while (n > BA_ROOT_MAX) {
    const bool warp = n <= WARP_LEVEL_MAX;
    const int K = warp ? 4 : PK;
    const uint32_t g2 = (uint32_t)((n + (uint64_t)K * BA_THREADS - 1) / ((uint64_t)K * BA_THREADS));
    char* nxt = cur + n * fe;
    if (warp) k_prod_fwd<C, true><<<g2, BA_THREADS, 0, s>>>(cur, (uint32_t)n, curp, nxt, ln_.others.p, K);
    else k_prod_fwd<C, false><<<g2, BA_THREADS, 0, s>>>(cur, (uint32_t)n, curp, nxt, nullptr, K);
    CKL();
    lv.push_back(Lvl{n, cur, curp, warp, K});
    curp += n * fe; cur = nxt; n = warp ? (uint64_t)g2 * (BA_THREADS / 32) : (uint64_t)g2 * BA_THREADS;
  }
  k_inv_root<C><<<1, RootCfg<C>::THREADS, 0, s>>>(cur, (uint32_t)n); CKL();
  for (int l = (int)lv.size() - 1; l >= 0; l--) {
    const Lvl& L = lv[l];
    const uint32_t g2 = (uint32_t)((L.n + (uint64_t)L.K * BA_THREADS - 1) / ((uint64_t)L.K * BA_THREADS));
    if (L.warp) k_prod_bwd<C, true><<<g2, BA_THREADS, 0, s>>>(L.v, (uint32_t)L.n, L.p, L.v + L.n * fe, ln_.others.p, L.K);
    else k_prod_bwd<C, false><<<g2, BA_THREADS, 0, s>>>(L.v, (uint32_t)L.n, L.p, L.v + L.n * fe, nullptr, L.K);
    CKL();
  }
  return B200MSM_OK;
}

// ---- batch-affine tree over a bucket range (a group of whole window slots), issued on one lane ----------------
// A lane = one CUDA stream + its own scratch.  Consecutive groups alternate between lanes so that the latency-bound
// tail of one group's round (product tree, root inversion) overlaps the throughput-bound kernels of the other group.
// off0 = sort offsets (absolute positions in sorted[]), cnt0 = bucket counts of the range,
// m0 = pairs in the range, maxcnt = largest bucket population in the range.  Writes buckets_g[0 .. nbg) as XYZZ.
template <class C>
int accumulate_batch_affine(b200msm_ctx* ctx, TreeLane& ln_, cudaStream_t s, uint32_t share, const void* d_bases, const void* xs, const uint32_t* off0, const uint32_t* cnt0, uint32_t nbg, uint64_t m0,
                            uint32_t maxcnt, void* buckets_g, uint32_t* rounds_out, uint64_t* adds_out) {
  const bool overlapped = share > 1;
  // s = the stream this group is issued on: the lane's own stream when several lanes overlap, the context's stream otherwise
  // (the lane then only lends its scratch; nothing in the lane is modified, so every error return leaves the context intact)
  const uint32_t* sorted = ctx->sorted.as<uint32_t>();
  // number of tree rounds: until the largest segment is <= 3 points (the finish kernel sums the rest serially)
  uint32_t R = 0;
  if (ctx->opt_tree_rounds >= 0) R = (uint32_t)ctx->opt_tree_rounds;
  else {
    // full tree: until the largest segment is <= 3 points.  Every round ends in a latency-bound tail (product tree + one
    // inversion, ~0.2 ms), so small problems stop earlier and let k_accum_finish sum up to 16 leftover points per bucket
    // serially (measured optimum: no round below 2^16 points, one at 2^16, 3 up to 2^18, 4 up to 2^19); buckets above 16 points still get
    // their rounds (need16), so skewed inputs keep the parallel tree.
    uint32_t full = 0, need16 = 0, mc = maxcnt;
    while (mc > 3) { if (mc > 16) need16++; mc = (mc + 1) >> 1; full++; }
    const uint64_t npts = ctx->cur_n;
    const uint32_t pref = npts < (1u << 16) ? 0u : npts < (1u << 17) ? 2u : npts <= (1u << 18) ? 3u : npts <= (1u << 19) ? 4u : 99u;
    R = std::min(full, std::max(pref, need16));
  }
  if (m0 < 2) R = 0;
  if (R > 30) R = 30;
  *rounds_out = std::max(*rounds_out, R);
  if (R == 0) {
    k_accum_finish<C, true><<<(nbg + 127) / 128, 128, 0, s>>>(d_bases, sorted, nullptr, 0, off0, nbg, buckets_g); CKL();
    MARK(T_FINISH);
    return B200MSM_OK;
  }
  // upper bounds of the slot counts per round: sum ceil(n/2) <= (sum n + #non-empty) / 2
  std::vector<uint64_t> U(R + 2); U[0] = m0;
  for (uint32_t r = 0; r <= R; r++) U[r + 1] = std::min<uint64_t>(U[r], (U[r] + std::min<uint64_t>(nbg, U[r]) + 1) / 2);
  // offsets for rounds 1..R in one fused scan (round 0 uses off0 directly)
  const size_t offstride = (size_t)nbg + 1;
  const uint32_t ntiles = (nbg + SCAN_TILE - 1) / SCAN_TILE;
  CK(ln_.offs.ensure(offstride * R * 4)); CK(ln_.tiles.ensure((size_t)(ntiles + 1) * R * 4));
  k_mscan_tiles<<<ntiles, SCAN_THREADS, 0, s>>>(cnt0, nbg, R, ln_.offs.as<uint32_t>(), ln_.tiles.as<uint32_t>(), ntiles); CKL();
  k_mscan_sums<<<R, 1024, 0, s>>>(ln_.tiles.as<uint32_t>(), ntiles); CKL();
  { dim3 g((nbg + 255) / 256, R); k_mscan_apply<<<g, 256, 0, s>>>(ln_.offs.as<uint32_t>(), nbg, ln_.tiles.as<uint32_t>(), ntiles); CKL(); }
  std::vector<const uint32_t*> off(R + 1); off[0] = off0;
  for (uint32_t r = 1; r <= R; r++) off[r] = ln_.offs.as<uint32_t>() + offstride * (r - 1);
  // bid arrays for rounds 1..R
  std::vector<uint32_t*> bid(R + 2, nullptr);
  { size_t tot = 0; for (uint32_t r = 1; r <= R; r++) tot += U[r];
    CK(ln_.bid.ensure(tot * 4 + 16));
    size_t at = 0; for (uint32_t r = 1; r <= R; r++) { bid[r] = ln_.bid.as<uint32_t>() + at; at += U[r]; } }
  k_fill_bid<<<(nbg + 255) / 256, 256, 0, s>>>(off[1], nbg, bid[1]); CKL();
  MARK(T_PLAN);
  const size_t fe = 4 * C::N, pt = 8 * C::N;
  const int BK = ctx->opt_ba_k > 0 ? ctx->opt_ba_k : (m0 >= (1u << 22) ? 16 : 8), PK = ctx->opt_pt_k;     // chain length per thread: measured 2^20: 8 -> 6.84, 12 -> 6.76, 16 -> 6.70 ms; 2^18: 2.70 / 2.69 / 2.80
  const uint64_t BA_TILE = (uint64_t)BK * BA_THREADS;
  CK(ln_.pa.ensure(U[1] * pt + 1024)); if (R > 1) CK(ln_.pb.ensure(U[2] * pt + 1024));
  // additions of round r: sum floor(n_r / 2) <= U[r] / 2 -- the arithmetic kernels run over these (dense) items, not over the output slots
  const uint64_t A0 = U[0] / 2;
  CK(ln_.prefix.ensure((A0 + BA_TILE) * fe + 16));
  // operand tables: one per round when they are all written up front (option "meta_upfront", off: measured 2^20 6.33 vs 6.25 ms, 2^18 2.61 vs 2.58 -- the later rounds' tables delay the first forward pass; a round's table depends only on the segment
  // offsets and the slot -> bucket maps, not on the points, so no k_tree_meta sits between a backward pass and the next forward pass), else one, reused
  const bool upfront = ctx->opt_meta_upfront != 0 && !ctx->prof;
  std::vector<uint64_t> ioff(R + 1, 0), coff(R + 1, 0);
  for (uint32_t r = 0; r < R; r++) { ioff[r + 1] = upfront ? ioff[r] + U[r] / 2 + 8 : 0; coff[r + 1] = upfront ? coff[r] + std::min<uint64_t>(nbg, U[r + 1]) + 8 : 0; }
  CK(ln_.meta.ensure((std::max<uint64_t>(ioff[R], A0) + BA_TILE) * 16 + 16)); CK(ln_.carry.ensure(((size_t)std::max<uint64_t>(coff[R], std::min<uint64_t>(nbg, U[1])) + 1) * 8));
  // product-tree level sizes for the largest round
  { uint64_t n1 = ((A0 + BA_TILE - 1) / BA_TILE + 1) * BA_THREADS;       // every level above is at least 4x smaller: 2*n1 bounds the sum
    CK(ln_.prod.ensure((2 * n1 + 4096) * fe)); CK(ln_.lvlprefix.ensure((2 * n1 + 4096) * fe)); CK(ln_.others.ensure((WARP_LEVEL_MAX / 4 + 4096) * fe)); }
#if defined(B200_EXPERIMENTS)
  const bool bt = ctx->opt_block_tree != 0 && !ctx->opt_fused;
#else
  const bool bt = false;
#endif
  if (bt) CK(ln_.btree.ensure(((A0 + BA_TILE - 1) / BA_TILE + 2) * (2 * BA_THREADS) * fe));      // 256 tree nodes per tile (tiles never outnumber round 0's at the shortest chain length)
  void* pin = nullptr; uint64_t yin = 0;
  const uint64_t ya = ((U[1] * fe + 255) / 256) * 256, yb = R > 1 ? ((U[2] * fe + 255) / 256) * 256 : 0;      // x array, then y array (see meta_load_point)
  auto launch_meta = [&](uint32_t r) -> int {
    TreeRound tr{off[r], off[r + 1], (r + 2 <= R) ? off[r + 2] : nullptr, bid[r + 1], (r + 2 <= R) ? bid[r + 2] : nullptr, nbg};
    const uint32_t nslots = (uint32_t)std::max<uint64_t>(U[r + 1], 1);
    uint4* items = ln_.meta.as<uint4>() + ioff[r]; uint2* carries = ln_.carry.as<uint2>() + coff[r];
    if (r == 0) k_tree_meta<true><<<(nslots + 255) / 256, 256, 0, s>>>(tr, sorted, items, carries);
    else k_tree_meta<false><<<(nslots + 255) / 256, 256, 0, s>>>(tr, nullptr, items, carries);
    CKL();
    return B200MSM_OK;
  };
  if (upfront) for (uint32_t r = 0; r < R; r++) { int rc_ = launch_meta(r); if (rc_) return rc_; }
  for (uint32_t r = 0; r < R; r++) {
    void* pout = (r & 1) ? ln_.pb.p : ln_.pa.p; const uint64_t yout = (r & 1) ? yb : ya;
    // chain length of this round: with the block-level tree a round whose tiles number <= BA_ROOT_MAX needs no product-tree launch at all,
    // so the chains are lengthened (up to 16) until they do
    int RK = BK;
    if (bt && ctx->opt_ba_k == 0) while (RK < 16 && (U[r] / 2 + (uint64_t)RK * BA_THREADS - 1) / ((uint64_t)RK * BA_THREADS) > BA_ROOT_MAX) RK *= 2;
    const uint64_t R_TILE = (uint64_t)RK * BA_THREADS;
    uint32_t grid = (uint32_t)((U[r] / 2 + R_TILE - 1) / R_TILE);
    if (grid == 0) grid = 1;
    char* prod = ln_.prod.as<char>(); char* lpre = ln_.lvlprefix.as<char>();
    const uint32_t pgrid = (ctx->opt_persist > 0 && overlapped) ? std::min<uint32_t>(grid, (uint32_t)ctx->opt_persist) : grid;   // persistent grid only when lanes overlap
    uint4* items = ln_.meta.as<uint4>() + ioff[r];
    uint2* carries = ln_.carry.as<uint2>() + coff[r];
    if (!upfront) { int rc_ = launch_meta(r); if (rc_) return rc_; }
#if defined(B200_EXPERIMENTS)
    if (ctx->opt_fused) {
      // the whole round in one launch, every CTA inverting its own tile's product (k_tree_round).  Small rounds use shorter chains so that the
      // round still spreads over the SMs (a tile is then cheaper than its inversion's latency, which is all such a round costs).
      MARK(T_TREE_FWD);
      int FK = BK;
      if (ctx->opt_ba_k == 0) { const uint64_t want = (U[r] / 2 + (uint64_t)BA_THREADS * ctx->opt_fused_tiles - 1) / ((uint64_t)BA_THREADS * ctx->opt_fused_tiles); FK = (int)std::min<uint64_t>(ctx->opt_fused_kmax, std::max<uint64_t>(2, want)); }
      uint32_t fg = (uint32_t)((U[r] / 2 + (uint64_t)FK * BA_THREADS - 1) / ((uint64_t)FK * BA_THREADS)); if (fg == 0) fg = 1;
      const uint32_t fpg = (ctx->opt_persist > 0 && overlapped) ? std::min<uint32_t>(fg, (uint32_t)ctx->opt_persist) : fg;
      bool piped = false;
      if constexpr (C::EXT == 1) { if (ctx->opt_fused == 3) {      // de-phased: CTAs sharing an SM start a third of a tile apart
        const uint32_t ppg = std::min<uint32_t>(fg, (uint32_t)ctx->opt_fused_grid), pdiv = std::max<uint32_t>(1, (uint32_t)ctx->opt_fused_grid / 3);
        if (r == 0) k_tree_round_pipe3<C, true><<<ppg, RP_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, d_bases, xs, 0, ln_.prefix.p, pout, yout, FK, fg, pdiv);
        else k_tree_round_pipe3<C, false><<<ppg, RP_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, pin, nullptr, yin, ln_.prefix.p, pout, yout, FK, fg, pdiv);
        piped = true; } }
      if constexpr (C::EXT == 1) { if (ctx->opt_fused == 2) {      // software-pipelined form: 4 compute warps + 1 inverting warp, 3 CTAs per SM
        const uint32_t ppg = std::min<uint32_t>(fg, (uint32_t)ctx->opt_fused_grid);
        if (r == 0) k_tree_round_pipe<C, true><<<ppg, RP_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, d_bases, xs, 0, ln_.prefix.p, pout, yout, FK, fg);
        else k_tree_round_pipe<C, false><<<ppg, RP_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, pin, nullptr, yin, ln_.prefix.p, pout, yout, FK, fg);
        piped = true; } }
      if (!piped) {
      if (r == 0) k_tree_round<C, true><<<fpg, BA_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, d_bases, 0, ln_.prefix.p, pout, yout, FK, fg);
      else k_tree_round<C, false><<<fpg, BA_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, pin, yin, ln_.prefix.p, pout, yout, FK, fg);
      }
      CKL(); MARK(r == 0 ? T_TREE_BWD0 : T_TREE_BWD);
    } else
#endif
    {
    const uint32_t fgrid = (ctx->opt_persist > 0 && overlapped) ? std::min<uint32_t>(grid, (uint32_t)(ctx->opt_persist_fwd > 0 ? ctx->opt_persist_fwd : ctx->opt_persist)) : grid;
#if defined(B200_EXPERIMENTS)
    if (bt) {      // first product-tree level inside the tree kernels: one value per tile goes up (block_upsweep / block_downsweep)
      void* gt = ln_.btree.p;
      if (r == 0) k_tree_fwd_bt<C, true><<<fgrid, BA_THREADS, 0, s>>>(items, off[r], off[r + 1], nbg, d_bases, xs, 0, ln_.prefix.p, prod, RK, grid, gt);
      else k_tree_fwd_bt<C, false><<<fgrid, BA_THREADS, 0, s>>>(items, off[r], off[r + 1], nbg, pin, nullptr, yin, ln_.prefix.p, prod, RK, grid, gt);
      CKL(); MARK(T_TREE_FWD);
      { int rc_ = product_tree_invert<C>(ctx, ln_, s, prod, lpre, (uint64_t)grid, PK); if (rc_) return rc_; }
      MARK(T_INV_TREE);
      if (r == 0) k_tree_bwd_bt<C, true><<<pgrid, BA_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, d_bases, 0, ln_.prefix.p, prod, pout, yout, RK, grid, gt);
      else k_tree_bwd_bt<C, false><<<pgrid, BA_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, pin, yin, ln_.prefix.p, prod, pout, yout, RK, grid, gt);
      CKL(); MARK(r == 0 ? T_TREE_BWD0 : T_TREE_BWD);
    } else
#endif
    {
    if (r == 0) k_tree_fwd<C, true><<<fgrid, BA_THREADS, 0, s>>>(items, off[r], off[r + 1], nbg, d_bases, xs, 0, ln_.prefix.p, prod, BK, grid);
    else k_tree_fwd<C, false><<<fgrid, BA_THREADS, 0, s>>>(items, off[r], off[r + 1], nbg, pin, nullptr, yin, ln_.prefix.p, prod, BK, grid);
    CKL(); MARK(T_TREE_FWD);
    { int rc_ = product_tree_invert<C>(ctx, ln_, s, prod, lpre, (uint64_t)grid * BA_THREADS, PK); if (rc_) return rc_; }
    MARK(T_INV_TREE);
    if (r == 0) k_tree_bwd<C, true><<<pgrid, BA_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, d_bases, 0, ln_.prefix.p, prod, pout, yout, BK, grid);
    else k_tree_bwd<C, false><<<pgrid, BA_THREADS, 0, s>>>(items, carries, off[r], off[r + 1], nbg, pin, yin, ln_.prefix.p, prod, pout, yout, BK, grid);
    CKL(); MARK(r == 0 ? T_TREE_BWD0 : T_TREE_BWD);
    }
    }
    pin = pout; yin = yout;
    *adds_out += U[r] - U[r + 1];
  }
  if (ctx->prof) {   // exact slot counts per round (the U[] are upper bounds): read the scan totals back, stats mode only
    for (uint32_t r = 1; r <= R; r++) CK(cudaMemcpyAsync(ctx->h_pinned + 1024 + r, off[r] + nbg, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    uint64_t prev = m0, exact = 0;
    for (uint32_t r = 1; r <= R; r++) { uint64_t cur = ctx->h_pinned[1024 + r]; exact += prev - cur; if (r == 1) ctx->adds_r0 += prev - cur; prev = cur; }
    *adds_out -= 0; ctx->adds_exact += exact;
  }
  k_accum_finish<C, false><<<(nbg + 127) / 128, 128, 0, s>>>(nullptr, nullptr, pin, yin, off[R], nbg, buckets_g); CKL();
  MARK(T_FINISH);
  return B200MSM_OK;
}

// in-place folding of `slots` consecutive bucket arrays of B buckets each (see k_fold)
template <class C>
int fold_slots(b200msm_ctx* ctx, cudaStream_t s, void* buckets_g, uint32_t slots, uint32_t B) {
  const uint32_t tail = std::min<uint32_t>(B, FOLD_TAIL);
  uint32_t sz = B;
  for (; sz > tail; sz >>= 1) {          // wide levels: one launch each (throughput-bound)
    uint32_t nblk = B / sz, live = 1; for (uint32_t k = 1; k < nblk; k <<= 1) live++;
    uint64_t threads = (uint64_t)live * (sz / 2) * slots;
    k_fold<C><<<(uint32_t)((threads + 127) / 128), 128, 0, s>>>(buckets_g, slots, B, sz); CKL();
  }
  if (tail >= 2) {                        // remaining levels of every live block of `tail` buckets: one launch
    uint32_t nblk = B / tail, live = 1; for (uint32_t k = 1; k < nblk; k <<= 1) live++;
#if defined(B200_EXPERIMENTS)
    bool quad = false;
    if constexpr (C::EXT == 1) quad = ctx->opt_fold_cluster == 2;
    if (quad) {                            // one cluster of 8 CTAs per live block, four lanes per addition (k_fold_tail_quad)
      if constexpr (C::EXT == 1) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(slots * live * FOLD_CLUSTER); cfg.blockDim = dim3(FOLD_QUAD_THREADS); cfg.stream = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = FOLD_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, k_fold_tail_quad<C>, buckets_g, B, tail, live)); CKL();
      }
    } else
#endif
    if (ctx->opt_fold_cluster) {          // one cluster of 8 CTAs per live block (k_fold_tail_cluster)
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(slots * live * FOLD_CLUSTER); cfg.blockDim = dim3(FOLD_CLUSTER_THREADS); cfg.stream = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = FOLD_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, k_fold_tail_cluster<C>, buckets_g, B, tail, live)); CKL();
    } else { k_fold_tail<C><<<slots * live, 256, 0, s>>>(buckets_g, B, tail, live); CKL(); }
  }
  return B200MSM_OK;
}

// ---- ordinary MSM with the sort pipelined per window group --------------------------------------------------------------------------
// The window slots are cut into groups BEFORE sorting (equal numbers of windows; scalars are taken to populate the windows evenly -- an
// empty group is skipped, an over-full one only takes longer).  Group g is counted, scanned and scattered on ITS lane's stream and its
// tree follows in stream order, so lane 0's accumulation starts after a quarter of the sort instead of all of it and the lanes' latency-bound
// tails (fold, read-back, host combination) come staggered instead of together.  Groups are issued from the TOP windows down (the host
// combination consumes them in that order); the last-issued group is the smallest, because its tail is the one nothing else overlaps.
// Region of group g in sorted[]: [n * w0, n * w1) -- a scalar has at most one pair per window (the extra slot shares the top window's).
template <class C>
int run_grouped(b200msm_ctx* ctx, const void* d_bases, const void* xs, const uint32_t* d_scal, const MsmPlan& pl, void* d_out) {
  cudaStream_t s = ctx->stream;
  const uint32_t n = pl.n, tb = 256, gb = (n + tb - 1) / tb;
  const uint32_t lanes = (uint32_t)std::max(1, std::min(ctx->opt_lanes, (int)MAX_LANES));
  const double per_pair = 8.0 * C::N * 0.75 + 4.0 * C::N * 0.75 + 6;
  const uint64_t budget_pairs = ctx->opt_group_pairs > 0 ? (uint64_t)ctx->opt_group_pairs : (uint64_t)std::max(1.0, 0.45 * ctx->mem_share * (double)ctx->total_mem / per_pair / lanes);
  uint32_t ngroups = std::max<uint32_t>(ctx->opt_groups > 0 ? (uint32_t)ctx->opt_groups : lanes, (uint32_t)(((uint64_t)n * pl.Wd + budget_pairs - 1) / budget_pairs));
  ngroups = std::min(ngroups, pl.Wd);
  // explicit plan (option "group_plan", tuning): window counts of the groups in ISSUE order (top group first), 5 bits each; used when they sum to Wd
  std::vector<uint32_t> plan;
  if (ctx->opt_group_plan > 0) { uint32_t sum = 0; for (int64_t v = ctx->opt_group_plan; v; v >>= 5) { plan.push_back((uint32_t)(v & 31)); sum += (uint32_t)(v & 31); }
    if (sum != pl.Wd || std::find(plan.begin(), plan.end(), 0u) != plan.end() || (uint64_t)n * pl.Wd > budget_pairs * plan.size()) plan.clear(); else ngroups = (uint32_t)plan.size(); }
  // window boundaries, bottom up: cut[0] = 0 .. cut[ngroups] = Wd; group ngroups-1 (top) is issued first, group 0 (bottom) last and is the smallest
  std::vector<uint32_t> cut(ngroups + 1, 0); cut[ngroups] = pl.Wd;
  if (!plan.empty()) { for (uint32_t g = ngroups; g-- > 1; ) cut[g] = cut[g + 1] - plan[ngroups - 1 - g]; }
  else { const double small = ngroups > 1 ? ctx->opt_group_small / 100.0 : 1.0, unit = (double)pl.Wd / ((double)ngroups - 1.0 + small);
    double acc = small * unit;
    for (uint32_t g = 1; g < ngroups; g++) { cut[g] = std::min(std::max<uint32_t>((uint32_t)(acc + 0.5), cut[g - 1] + 1), pl.Wd - (ngroups - g)); acc += unit; } }
  const uint32_t per = pl.logB + 1, npts = pl.W * per;
  const size_t fbytes = (size_t)npts * 16 * C::N;
  CK(ctx->wsum.ensure(fbytes));
  if (ctx->h_folded_cap < fbytes + 4096) { if (ctx->h_folded) cudaFreeHost(ctx->h_folded); ctx->h_folded = nullptr; ctx->h_folded_cap = 0;
    CK(cudaMallocHost(&ctx->h_folded, fbytes + 4096)); ctx->h_folded_cap = fbytes + 4096; }
  while (ctx->gev.size() < 2 * (size_t)ngroups) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (ctx->blocking_waits ? cudaEventBlockingSync : 0))); ctx->gev.push_back(e); }
  CK(ctx->buckets.ensure((size_t)pl.W * pl.B * 16 * C::N));
  CK(ctx->misc.ensure(512 * 4));
  CK(cudaMemsetAsync(ctx->misc.p, 0, 512 * 4, s));
  CK(cudaEventRecord(ctx->ev_sorted, s));                      // inputs staged, scratch zeroed: the lanes may start
  int rc;
  for (uint32_t l = 0; l < lanes; l++) { rc = lane_init(ctx, ctx->lane[l]); if (rc) return rc; CK(cudaStreamWaitEvent(ctx->lane[l].stream, ctx->ev_sorted, 0)); }
  // ---- phase 1: every group's sort on its lane (all issued up front; the lanes run them concurrently)
  struct Grp { uint32_t w0, w1, s1, b0, nbg; uint32_t* offs; };
  std::vector<Grp> grp(ngroups);
  for (uint32_t gi = 0; gi < ngroups; gi++) {
    const uint32_t g = ngroups - 1 - gi;
    TreeLane& ln = ctx->lane[gi % lanes]; cudaStream_t ls = ln.stream;
    Grp& G = grp[gi];
    G.w0 = cut[g]; G.w1 = cut[g + 1]; G.s1 = (G.w1 == pl.Wd) ? pl.W : G.w1;           // slots [w0, s1): the top group owns the extra slot
    G.b0 = G.w0 * pl.B; G.nbg = (G.s1 - G.w0) * pl.B;
    G.offs = ctx->offsets.as<uint32_t>() + g;                                           // group g's nbg + 1 offsets live at [b0 + g, b0 + g + nbg] (g counted bottom-up): indexable by GLOBAL bucket id, no overlap between groups
    uint32_t* cnt = ctx->counts.as<uint32_t>() + G.b0;
    CK(cudaMemsetAsync(cnt, 0, (size_t)G.nbg * 4, ls));
    k_digits<false><<<gb, tb, 0, ls>>>(d_scal, pl, ctx->counts.as<uint32_t>(), nullptr, ctx->ranks.as<uint32_t>(), nullptr, G.w0, G.w1); CKL();
    { const uint32_t ntiles = (G.nbg + SCAN_TILE - 1) / SCAN_TILE;
      CK(ln.tiles.ensure((size_t)(ntiles + 1) * 64 * 4)); }                              // (also serves the tree's multi-round scan, which sizes it again)
    rc = exclusive_scan(ctx, ls, ln.tiles, cnt, G.offs + G.b0, G.nbg, n * G.w0); if (rc) return rc;
    { dim3 gd((pl.B + 255) / 256, G.s1 - G.w0); k_window_max<<<gd, 256, 0, ls>>>(cnt, pl.B, ctx->misc.as<uint32_t>() + G.w0); CKL(); }
    CK(cudaMemcpyAsync(ctx->h_pinned + G.w0, ctx->misc.as<uint32_t>() + G.w0, (G.s1 - G.w0) * 4, cudaMemcpyDeviceToHost, ls));
    CK(cudaMemcpyAsync(ctx->h_pinned + 512 + gi, G.offs + G.b0 + G.nbg, 4, cudaMemcpyDeviceToHost, ls));      // base + pairs of the group
    CK(cudaEventRecord(ctx->gev[ngroups + gi], ls));
    k_digits<true><<<gb, tb, 0, ls>>>(d_scal, pl, nullptr, G.offs, ctx->ranks.as<uint32_t>(), ctx->sorted.as<uint32_t>(), G.w0, G.w1); CKL();
    if (ctx->bases_pending) CK(cudaStreamWaitEvent(ls, ctx->ev_bases, 0));
  }
  // ---- phase 2: as each group's counts arrive, its tree, fold and read-back follow on the same lane (one issuing thread per lane)
  ctx->adds_r0 = 0; ctx->adds_exact = 0; ctx->cur_n = n;
  std::vector<char> live(ngroups, 0); std::vector<uint32_t> g_rounds(ngroups, 0); std::vector<uint64_t> g_adds(ngroups, 0);
  GroupIssue gi_(ctx, ngroups, lanes);
  for (uint32_t gi = 0; gi < ngroups; gi++) {
    auto job = [&, gi]() -> int {
      const Grp& G = grp[gi];
      TreeLane& ln = ctx->lane[gi % lanes]; cudaStream_t ls = ln.stream;
      CK(cudaEventSynchronize(ctx->gev[ngroups + gi]));
      uint32_t mc = 0; for (uint32_t w = G.w0; w < G.s1; w++) mc = std::max(mc, ctx->h_pinned[w]);
      const uint64_t m0 = (uint64_t)ctx->h_pinned[512 + gi] - (uint64_t)n * G.w0;
      const size_t o = (size_t)G.w0 * per * 16 * C::N; const uint32_t np = (G.s1 - G.w0) * per;
      if (m0 == 0) { memset(reinterpret_cast<char*>(ctx->h_folded) + o, 0, (size_t)np * 16 * C::N); return B200MSM_OK; }      // every digit of these windows is zero: all their folded entries are infinity (zz = 0)
      live[gi] = 1;
      char* bg = ctx->buckets.as<char>() + (size_t)G.b0 * 16 * C::N;
      int rc2 = accumulate_batch_affine<C>(ctx, ln, ls, lanes, d_bases, xs, G.offs + G.b0, ctx->counts.as<uint32_t>() + G.b0, G.nbg, m0, mc, bg, &g_rounds[gi], &g_adds[gi]);
      if (!rc2) rc2 = fold_slots<C>(ctx, ls, bg, G.s1 - G.w0, pl.B);
      if (rc2) return rc2;
      k_gather_folded<C><<<(np + 127) / 128, 128, 0, ls>>>(bg, G.s1 - G.w0, pl.B, pl.logB, ctx->wsum.as<char>() + o); CKL();
      CK(cudaMemcpyAsync(reinterpret_cast<char*>(ctx->h_folded) + o, ctx->wsum.as<char>() + o, (size_t)np * 16 * C::N, cudaMemcpyDeviceToHost, ls));
      CK(cudaEventRecord(ctx->gev[gi], ls));
      return B200MSM_OK;
    };
    rc = gi_.run(gi, gi % lanes, job); if (rc) return rc;
  }
  // ---- window combination on the host, group by group from the top (host_ec.h)
  using HF = typename HostField<C>::type; const HF f = HostField<C>::make();
  b200host::Combiner<HF> cb; cb.begin(f, pl.W, pl.Wd, pl.c0, pl.rem, pl.logB);
  float host_ms = 0;
  for (uint32_t gi = 0; gi < ngroups; gi++) {
    rc = gi_.wait(gi); if (rc) return rc;
    if (live[gi]) CK(cudaEventSynchronize(ctx->gev[gi]));
    auto t0 = std::chrono::steady_clock::now();
    cb.feed(reinterpret_cast<const b200host::XYZZ<HF::W>*>(ctx->h_folded), grp[gi].w0, grp[gi].s1);
    host_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  auto t0 = std::chrono::steady_clock::now();
  uint64_t* res = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(ctx->h_folded) + fbytes);
  cb.finish(res);
  ctx->host_combine_ms = host_ms + std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  for (uint32_t l = 0; l < lanes; l++) { CK(cudaEventRecord(ctx->lane[l].done, ctx->lane[l].stream)); CK(cudaStreamWaitEvent(s, ctx->lane[l].done, 0)); }
  CK(cudaMemcpyAsync(d_out, res, 12 * C::N, cudaMemcpyHostToDevice, s));
  return B200MSM_OK;
}

// Core pipeline: d_bases (affine Montgomery, device), d_scal (canonical 8-word scalars, device), result -> d_out (device, 3*n8 bytes)
template <class C>
int run_pipeline(b200msm_ctx* ctx, const void* d_bases, const uint32_t* d_scal, uint64_t n64, uint32_t nbits, void* d_out, b200msm_stats* st, const Precomp* pre) {
  cudaStream_t s = ctx->stream;
  const uint32_t n = (uint32_t)n64;
  MsmPlan pl;
  pl.n = n; pl.nbits = nbits; pl.pre_stride = 0;
  if (pre) {   // the plan the table was built for; all windows share one bucket array (two slots when the unsigned last digit can reach 2B)
    pl.Wd = pre->Wd; pl.c0 = pre->c0; pl.rem = pre->rem; pl.c = pl.c0 + (pl.rem ? 1 : 0); pl.B = 1u << (pl.c - 1); pl.logB = pl.c - 1;
    pl.W = pl.rem == 0 ? 2 : 1; pl.pre_stride = pre->stride;
  } else
  { // window plan: target width ct, then equalise: Wd windows whose widths differ by at most one bit
    uint32_t ct = ctx->opt_window_bits > 0 ? std::min<uint32_t>((uint32_t)ctx->opt_window_bits, std::min<uint32_t>(nbits, 24)) : auto_window_bits(n, nbits);
    pl.Wd = (nbits + ct - 1) / ct; pl.c0 = nbits / pl.Wd; pl.rem = nbits - pl.c0 * pl.Wd;
    pl.c = pl.c0 + (pl.rem ? 1 : 0); pl.B = 1u << (pl.c - 1); pl.logB = pl.c - 1;
    pl.W = pl.Wd + (pl.rem == 0 ? 1 : 0); }
  const uint32_t nb = pl.W * pl.B;
  const void* xs = pre ? nullptr : ctx->cur_xs;
  if (pre && (uint64_t)pre->stride * pl.Wd >= (1ull << 31)) { ctx->err = "window table too large for 31-bit point indices"; return B200MSM_E_UNSUPPORTED; }
  if ((uint64_t)pl.W * pl.B > (1ull << 31) || (uint64_t)n * std::max(pl.W, pl.Wd) >= (1ull << 32)) { ctx->err = "problem too large for 32-bit pair indices"; return B200MSM_E_UNSUPPORTED; }
  if (pl.W > 400) { ctx->err = "too many windows"; return B200MSM_E_UNSUPPORTED; }

  if (st) CK(cudaEventRecord(ctx->ev[1], s));
  CK(ctx->counts.ensure((size_t)nb * 4)); CK(ctx->offsets.ensure((size_t)(nb + 1 + 512) * 4));
  CK(ctx->sorted.ensure((size_t)n * std::max(pl.W, pl.Wd) * 4 + 16));
  CK(cudaMemsetAsync(ctx->counts.p, 0, (size_t)nb * 4, s));
  const uint32_t tb = 256, gb = (n + tb - 1) / tb;
  CK(ctx->ranks.ensure((size_t)n * pl.Wd * 4 + 16));
  // ---- pipelined form (large ordinary MSMs, several lanes): every window group is sorted on its own lane's stream, see run_grouped
  if (!pre && !ctx->prof && ctx->opt_sort_groups && ctx->opt_accumulate != 1 && ctx->opt_combine == 0 && ctx->opt_lanes > 1 && n >= (1u << 18))
    return run_grouped<C>(ctx, d_bases, xs, d_scal, pl, d_out);
  k_digits<false><<<gb, tb, 0, s>>>(d_scal, pl, ctx->counts.as<uint32_t>(), nullptr, ctx->ranks.as<uint32_t>(), nullptr, 0u, pl.Wd); CKL();
  int rc = exclusive_scan(ctx, s, ctx->tiles, ctx->counts.as<uint32_t>(), ctx->offsets.as<uint32_t>(), nb, 0u);
  if (rc) return rc;
  // ---- per-slot pair counts and largest bucket populations go back to the host while the scatter runs
  CK(ctx->misc.ensure(512 * 4));
  CK(cudaMemsetAsync(ctx->misc.p, 0, 512 * 4, s));
  // planning granules: whole slots normally; with a window table (one or two slots only) the slot is cut into up to 64 bucket ranges
  uint32_t Bg = pl.B;
  if (pre) {   // sub-slots (see the accumulate block): 8 per slot by default, more when a sub-slot's tree scratch would exceed the memory budget
    const double per_pair = 8.0 * C::N * 0.75 + 4.0 * C::N * 0.75 + 6;
    const double budget = ctx->opt_group_pairs > 0 ? (double)ctx->opt_group_pairs : 0.45 * ctx->mem_share * (double)ctx->total_mem / per_pair / MAX_LANES;
    uint32_t S = ctx->opt_subslots > 0 ? (uint32_t)ctx->opt_subslots : 4;      // measured at 2^20, c = 20: 4 -> 5.86 ms, 8 -> 5.97, 16 -> 5.96, 32 -> 6.23
    while (S < 256 && (double)n * pl.Wd / S > budget) S *= 2;
    while (S > 1 && pl.B / S < 64) S /= 2;
    if (pl.B / S >= 1) Bg = pl.B / S;
  }
  const uint32_t G = nb / Bg;
  { dim3 g((Bg + 255) / 256, G); k_window_max<<<g, 256, 0, s>>>(ctx->counts.as<uint32_t>(), Bg, ctx->misc.as<uint32_t>()); CKL(); }
  CK(cudaMemcpyAsync(ctx->h_pinned, ctx->misc.p, G * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpy2DAsync(ctx->h_pinned + 512, 4, ctx->offsets.as<uint32_t>(), (size_t)Bg * 4, 4, G + 1, cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(ctx->ev_plan, s));
  k_digits<true><<<gb, tb, 0, s>>>(d_scal, pl, nullptr, ctx->offsets.as<uint32_t>(), ctx->ranks.as<uint32_t>(), ctx->sorted.as<uint32_t>(), 0u, pl.Wd); CKL();
  if (st) CK(cudaEventRecord(ctx->ev[2], s));
  CK(cudaEventRecord(ctx->ev_sorted, s));
  CK(cudaEventSynchronize(ctx->ev_plan));
  MARK(T_SORT);
  // ---- accumulate: bucket sums in XYZZ, then fold every slot in place
  if (ctx->bases_pending) CK(cudaStreamWaitEvent(s, ctx->ev_bases, 0));
  CK(ctx->buckets.ensure((size_t)nb * 16 * C::N));
  int mode = ctx->opt_accumulate;
  if (mode == 0) mode = 2;
  uint32_t rounds = 0; uint64_t adds = 0;
  ctx->adds_r0 = 0; ctx->adds_exact = 0; ctx->cur_n = n;
  if (pre) {
    // ---- window-table form: the single bucket array is cut into S sub-slots of Bs = Bg buckets (a power of two).  Each sub-slot is a
    // group: its tree runs on one lane, then it is folded on its own (k_fold sees it as a slot of Bs buckets) while the other lanes still
    // accumulate, so the latency-bound fold levels are hidden.  With b = s*Bs + l:  sum_b (b+1) T[b] = sum_s [ F_s + s*Bs * T_s ],
    // F_s = folded value of sub-slot s, T_s = its plain sum = folded entry 0; the host adds these few terms (combine_subslots).
    const uint64_t mtot = ctx->h_pinned[512 + G];
    if (mtot == 0) {
      uint32_t z[3 * C::N]; memset(z, 0, sizeof z); for (int i = 0; i < C::N; i++) z[C::N + i] = C::one(i);
      memcpy(ctx->h_pinned + 1100, z, sizeof z);
      CK(cudaMemcpyAsync(d_out, ctx->h_pinned + 1100, sizeof z, cudaMemcpyHostToDevice, s));
      if (st) { CK(cudaEventRecord(ctx->ev[3], s)); CK(cudaEventRecord(ctx->ev[4], s)); CK(cudaEventRecord(ctx->ev[5], s));
                st->n = n; st->window_bits = pl.c; st->windows = pl.Wd; st->reserved = pl.W; st->buckets_per_window = pl.B; }
      return B200MSM_OK;
    }
    uint32_t lanes = ctx->prof ? 1u : (uint32_t)std::max(1, std::min(ctx->opt_lanes, (int)MAX_LANES));
    if (mtot < (1u << 16)) lanes = 1;
    const bool host_tail = ctx->opt_combine == 0;
    uint32_t logBs = 0; while ((1u << logBs) < Bg) logBs++;
    const uint32_t per = logBs + 1;
    const size_t fbytes = (size_t)G * per * 16 * C::N;
    if (host_tail) {
      CK(ctx->wsum.ensure(fbytes));
      if (ctx->h_folded_cap < fbytes + 4096) { if (ctx->h_folded) cudaFreeHost(ctx->h_folded); ctx->h_folded = nullptr; ctx->h_folded_cap = 0;
        CK(cudaMallocHost(&ctx->h_folded, fbytes + 4096)); ctx->h_folded_cap = fbytes + 4096; }
      while (ctx->gev.size() < G) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (ctx->blocking_waits ? cudaEventBlockingSync : 0))); ctx->gev.push_back(e); }
    }
    for (uint32_t l = 0; l < lanes; l++) { rc = lane_init(ctx, ctx->lane[l]); if (rc) return rc; if (lanes > 1) CK(cudaStreamWaitEvent(ctx->lane[l].stream, ctx->ev_sorted, 0)); }
    // groups = contiguous runs of sub-slots with (nearly) equal pair counts, one tree each; at least one per lane
    const double per_pair = 8.0 * C::N * 0.75 + 4.0 * C::N * 0.75 + 6;
    const uint64_t budget_pairs = ctx->opt_group_pairs > 0 ? (uint64_t)ctx->opt_group_pairs : (uint64_t)std::max(1.0, 0.45 * ctx->mem_share * (double)ctx->total_mem / per_pair / lanes);
    uint32_t ngroups = std::min<uint32_t>(G, std::max<uint32_t>(lanes, (uint32_t)((mtot + budget_pairs - 1) / budget_pairs)));
    std::vector<uint32_t> cut(ngroups + 1, 0); cut[ngroups] = G;
    { uint32_t w = 0; for (uint32_t g = 1; g < ngroups; g++) { uint64_t target = mtot * g / ngroups;
        while (w < G && ctx->h_pinned[512 + w] < target) w++;
        cut[g] = std::min(std::max(w, cut[g - 1] + 1), G - (ngroups - g)); } }
    std::vector<uint32_t> g_rounds(ngroups, 0); std::vector<uint64_t> g_adds(ngroups, 0);
    GroupIssue gi_(ctx, ngroups, lanes);
    for (uint32_t g = 0; g < ngroups; g++) {
      auto job = [&, g]() -> int {
        TreeLane& ln = ctx->lane[g % lanes];
        const uint32_t g0 = cut[g], g1 = cut[g + 1];
        uint32_t mc = 0; for (uint32_t k = g0; k < g1; k++) mc = std::max(mc, ctx->h_pinned[k]);
        const uint64_t m0 = ctx->h_pinned[512 + g1] - ctx->h_pinned[512 + g0];
        const uint32_t b0 = g0 * Bg, nbg = (g1 - g0) * Bg;
        cudaStream_t ls = lanes == 1 ? s : ln.stream;
        char* bg = ctx->buckets.as<char>() + (size_t)b0 * 16 * C::N;
        int rc2 = accumulate_batch_affine<C>(ctx, ln, ls, lanes, d_bases, nullptr, ctx->offsets.as<uint32_t>() + b0, ctx->counts.as<uint32_t>() + b0, nbg, m0, mc, bg, &g_rounds[g], &g_adds[g]);
        if (!rc2 && host_tail) {
          if (st && g + 1 == ngroups) CK(cudaEventRecord(ctx->ev[3], s));
          rc2 = fold_slots<C>(ctx, ls, bg, g1 - g0, Bg);
          if (!rc2) {
            const uint32_t np = (g1 - g0) * per; const size_t o = (size_t)g0 * per * 16 * C::N;
            k_gather_folded<C><<<(np + 127) / 128, 128, 0, ls>>>(bg, g1 - g0, Bg, logBs, ctx->wsum.as<char>() + o); CKL();
            CK(cudaMemcpyAsync(reinterpret_cast<char*>(ctx->h_folded) + o, ctx->wsum.as<char>() + o, (size_t)np * 16 * C::N, cudaMemcpyDeviceToHost, ls));
            CK(cudaEventRecord(ctx->gev[g], ls));
          }
        }
        return rc2;
      };
      rc = gi_.run(g, g % lanes, job); if (rc) return rc;
    }
    if (!host_tail) { rc = gi_.wait_all(); if (rc) return rc; }
    auto merge_stats = [&]() { for (uint32_t g = 0; g < ngroups; g++) { rounds = std::max(rounds, g_rounds[g]); adds += g_adds[g]; } };
    if (host_tail) {
      if (st) { MARK(T_FOLD); CK(cudaEventRecord(ctx->ev[4], s)); }
      using HF = typename HostField<C>::type; const HF f = HostField<C>::make();
      b200host::SubslotCombiner<HF> cb; cb.begin(f, G, logBs);
      float host_ms = 0;
      for (uint32_t g = 0; g < ngroups; g++) {          // each group's sub-slots are reduced as soon as they arrive, while the other lanes still run
        rc = gi_.wait(g); if (rc) return rc;
        CK(cudaEventSynchronize(ctx->gev[g]));
        auto t0 = std::chrono::steady_clock::now();
        cb.feed(reinterpret_cast<const b200host::XYZZ<HF::W>*>(ctx->h_folded), cut[g], cut[g + 1]);
        host_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
      uint64_t* res = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(ctx->h_folded) + fbytes);
      cb.finish(res);
      ctx->host_combine_ms = host_ms; merge_stats();
      if (lanes > 1) for (uint32_t l = 0; l < lanes; l++) { CK(cudaEventRecord(ctx->lane[l].done, ctx->lane[l].stream)); CK(cudaStreamWaitEvent(s, ctx->lane[l].done, 0)); }
      CK(cudaMemcpyAsync(d_out, res, 12 * C::N, cudaMemcpyHostToDevice, s));
      MARK(T_HORNER);
    } else {   // device chain (cross-check form): fold the whole array as pl.W slots of B buckets
      merge_stats();
      if (lanes > 1) for (uint32_t l = 0; l < lanes; l++) { CK(cudaEventRecord(ctx->lane[l].done, ctx->lane[l].stream)); CK(cudaStreamWaitEvent(s, ctx->lane[l].done, 0)); }
      if (st) CK(cudaEventRecord(ctx->ev[3], s));
      rc = fold_slots<C>(ctx, s, ctx->buckets.p, pl.W, pl.B); if (rc) return rc;
      MARK(T_FOLD);
      if (st) CK(cudaEventRecord(ctx->ev[4], s));
      CK(ctx->wsum.ensure((size_t)pl.W * 16 * C::N));
      k_window_sums<C><<<(pl.W + 31) / 32, 32, 0, s>>>(ctx->buckets.p, pl.W, 1, pl.B, pl.logB, ctx->wsum.p); CKL();
      MARK(T_WSUM);
      k_horner<C><<<1, 32, 0, s>>>(ctx->wsum.p, pl.W, 1, pl.c0, 0, d_out); CKL();
      MARK(T_HORNER);
    }
  } else if (mode == 2) {
    // groups of whole slots: at least `lanes` of them (overlap), more if the tree scratch would not fit in device memory
    const uint64_t mtot = ctx->h_pinned[512 + pl.W];
    uint32_t lanes = ctx->prof ? 1u : (uint32_t)std::max(1, std::min(ctx->opt_lanes, (int)MAX_LANES));
    if (mtot < (1u << 16)) lanes = 1;
    const double per_pair = 8.0 * C::N * 0.75 + 4.0 * C::N * 0.75 + 6;      // points (pa+pb) + prefix/products + bid, per input pair
    const uint64_t budget_pairs = ctx->opt_group_pairs > 0 ? (uint64_t)ctx->opt_group_pairs : (uint64_t)std::max(1.0, 0.45 * ctx->mem_share * (double)ctx->total_mem / per_pair / lanes);
    uint32_t ngroups = std::max<uint32_t>(lanes, (uint32_t)((mtot + budget_pairs - 1) / budget_pairs));
    // slots above the highest non-empty one are skipped altogether (short scalars in wide containers, e.g. GLV halves)
    uint32_t Wuse = pl.W; while (Wuse > 0 && ctx->h_pinned[512 + Wuse] == ctx->h_pinned[512 + Wuse - 1]) Wuse--;
    if (Wuse == 0) {                                              // every digit is zero: the sum is the point at infinity
      uint32_t z[3 * C::N]; memset(z, 0, sizeof z); for (int i = 0; i < C::N; i++) z[C::N + i] = C::one(i);
      memcpy(ctx->h_pinned + 1100, z, sizeof z);
      CK(cudaMemcpyAsync(d_out, ctx->h_pinned + 1100, sizeof z, cudaMemcpyHostToDevice, s));
      if (st) { CK(cudaEventRecord(ctx->ev[3], s)); CK(cudaEventRecord(ctx->ev[4], s)); CK(cudaEventRecord(ctx->ev[5], s));
                st->n = n; st->window_bits = pl.c; st->windows = pl.Wd; st->reserved = pl.W; st->buckets_per_window = pl.B; }
      return B200MSM_OK;
    }
    const uint32_t wlim = Wuse > pl.Wd ? pl.Wd : Wuse;            // the extra slot stays in the group of the last window
    ngroups = std::min(ngroups, wlim);
    // slot boundaries with (nearly) equal pair counts
    std::vector<uint32_t> cut(ngroups + 1, 0); cut[ngroups] = Wuse;
    { uint32_t w = 0; for (uint32_t g = 1; g < ngroups; g++) { uint64_t target = mtot * g / ngroups;
        while (w < Wuse && ctx->h_pinned[512 + w] < target) w++;
        cut[g] = std::min(std::max(w, cut[g - 1] + 1), wlim - (ngroups - g)); } }
    for (uint32_t l = 0; l < lanes; l++) { rc = lane_init(ctx, ctx->lane[l]); if (rc) return rc; if (lanes > 1) { CK(cudaStreamWaitEvent(ctx->lane[l].stream, ctx->ev_sorted, 0)); if (ctx->bases_pending) CK(cudaStreamWaitEvent(ctx->lane[l].stream, ctx->ev_bases, 0)); } }
    // host staging for the folded entries of every slot, and one event per group
    const uint32_t per = pl.logB + 1, npts = pl.W * per;
    const size_t fbytes = (size_t)npts * 16 * C::N;
    const bool host_tail = ctx->opt_combine == 0;
    if (host_tail) {
      CK(ctx->wsum.ensure(fbytes));
      if (ctx->h_folded_cap < fbytes + 4096) { if (ctx->h_folded) cudaFreeHost(ctx->h_folded); ctx->h_folded = nullptr; ctx->h_folded_cap = 0;
        CK(cudaMallocHost(&ctx->h_folded, fbytes + 4096)); ctx->h_folded_cap = fbytes + 4096; }
      while (ctx->gev.size() < ngroups) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (ctx->blocking_waits ? cudaEventBlockingSync : 0))); ctx->gev.push_back(e); }
    }
    // groups are issued from the TOP windows down: lane 0 (highest priority) owns the top group, so its folded points
    // reach the host first and the serial window combination runs while the GPU is still busy with the lower windows
    std::vector<uint32_t> g_rounds(ngroups, 0); std::vector<uint64_t> g_adds(ngroups, 0);
    GroupIssue gi_(ctx, ngroups, lanes);
    for (uint32_t gi = 0; gi < ngroups; gi++) {
      auto job = [&, gi]() -> int {
        const uint32_t g = ngroups - 1 - gi;
        TreeLane& ln = ctx->lane[gi % lanes];
        uint32_t w0 = cut[g], w1 = cut[g + 1];
        uint32_t mc = 0; for (uint32_t w = w0; w < w1; w++) mc = std::max(mc, ctx->h_pinned[w]);
        uint64_t m0 = ctx->h_pinned[512 + w1] - ctx->h_pinned[512 + w0];
        uint32_t b0 = w0 * pl.B, nbg = (w1 - w0) * pl.B;
        cudaStream_t ls = lanes == 1 ? s : ln.stream;
        char* bg = ctx->buckets.as<char>() + (size_t)b0 * 16 * C::N;
        int rc2 = accumulate_batch_affine<C>(ctx, ln, ls, lanes, d_bases, xs, ctx->offsets.as<uint32_t>() + b0, ctx->counts.as<uint32_t>() + b0, nbg, m0, mc, bg, &g_rounds[gi], &g_adds[gi]);
        if (!rc2 && (lanes > 1 || host_tail)) {
          if (st && gi + 1 == ngroups) CK(cudaEventRecord(ctx->ev[3], s));          // stats mode is single-lane: accumulate ends here
          rc2 = fold_slots<C>(ctx, ls, bg, w1 - w0, pl.B); }
        if (!rc2 && host_tail) {
          const uint32_t np = (w1 - w0) * per; const size_t o = (size_t)w0 * per * 16 * C::N;
          k_gather_folded<C><<<(np + 127) / 128, 128, 0, ls>>>(bg, w1 - w0, pl.B, pl.logB, ctx->wsum.as<char>() + o); CKL();
          CK(cudaMemcpyAsync(reinterpret_cast<char*>(ctx->h_folded) + o, ctx->wsum.as<char>() + o, (size_t)np * 16 * C::N, cudaMemcpyDeviceToHost, ls));
          CK(cudaEventRecord(ctx->gev[gi], ls));
        }
        return rc2;
      };
      rc = gi_.run(gi, gi % lanes, job); if (rc) return rc;
    }
    if (!host_tail || st) { rc = gi_.wait_all(); if (rc) return rc; }
    auto merge_stats = [&]() { for (uint32_t g = 0; g < ngroups; g++) { rounds = std::max(rounds, g_rounds[g]); adds += g_adds[g]; } };
    if (st) { if (!host_tail) CK(cudaEventRecord(ctx->ev[3], s)); else { MARK(T_FOLD); } CK(cudaEventRecord(ctx->ev[4], s)); }
    if (host_tail) {
      using HF = typename HostField<C>::type; const HF f = HostField<C>::make();
      b200host::Combiner<HF> cb; cb.begin(f, pl.W, pl.Wd, pl.c0, pl.rem, pl.logB);
      float host_ms = 0;
      for (uint32_t gi = 0; gi < ngroups; gi++) {
        const uint32_t g = ngroups - 1 - gi;
        rc = gi_.wait(gi); if (rc) return rc;
        CK(cudaEventSynchronize(ctx->gev[gi]));
        auto t0 = std::chrono::steady_clock::now();
        cb.feed(reinterpret_cast<const b200host::XYZZ<HF::W>*>(ctx->h_folded), cut[g], cut[g + 1]);
        host_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
      auto t0 = std::chrono::steady_clock::now();
      uint64_t* res = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(ctx->h_folded) + fbytes);   // 3*n8 bytes in the pinned tail
      cb.finish(res);
      ctx->host_combine_ms = host_ms + std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      merge_stats();
      if (lanes > 1) for (uint32_t l = 0; l < lanes; l++) { CK(cudaEventRecord(ctx->lane[l].done, ctx->lane[l].stream)); CK(cudaStreamWaitEvent(s, ctx->lane[l].done, 0)); }
      CK(cudaMemcpyAsync(d_out, res, 12 * C::N, cudaMemcpyHostToDevice, s));
      MARK(T_HORNER);
    } else {
      merge_stats();
      if (lanes > 1) for (uint32_t l = 0; l < lanes; l++) { CK(cudaEventRecord(ctx->lane[l].done, ctx->lane[l].stream)); CK(cudaStreamWaitEvent(s, ctx->lane[l].done, 0)); }
      if (Wuse < pl.W) CK(cudaMemsetAsync(ctx->buckets.as<char>() + (size_t)Wuse * pl.B * 16 * C::N, 0, (size_t)(pl.W - Wuse) * pl.B * 16 * C::N, s));   // skipped slots = infinity (zz = 0)
      if (lanes == 1) { rc = fold_slots<C>(ctx, s, ctx->buckets.p, pl.W, pl.B); if (rc) return rc; }
      MARK(T_FOLD);
      if (st) CK(cudaEventRecord(ctx->ev[4], s));
      CK(ctx->wsum.ensure((size_t)pl.W * 16 * C::N));
      k_window_sums<C><<<(pl.W + 31) / 32, 32, 0, s>>>(ctx->buckets.p, pl.W, pl.Wd, pl.B, pl.logB, ctx->wsum.p); CKL();
      MARK(T_WSUM);
      k_horner<C><<<1, 32, 0, s>>>(ctx->wsum.p, pl.W, pl.Wd, pl.c0, pl.rem, d_out); CKL();
      MARK(T_HORNER);
    }
  } else {
    // serial accumulate (one thread per bucket): small problems / cross-check; device-side combination
    k_accum_finish<C, true><<<(nb + 127) / 128, 128, 0, s>>>(d_bases, ctx->sorted.as<uint32_t>(), nullptr, 0, ctx->offsets.as<uint32_t>(), nb, ctx->buckets.p); CKL();
    if (st) CK(cudaEventRecord(ctx->ev[3], s));
    rc = fold_slots<C>(ctx, s, ctx->buckets.p, pl.W, pl.B); if (rc) return rc;
    MARK(T_FOLD);
    if (st) CK(cudaEventRecord(ctx->ev[4], s));
    CK(ctx->wsum.ensure((size_t)pl.W * 16 * C::N));
    k_window_sums<C><<<(pl.W + 31) / 32, 32, 0, s>>>(ctx->buckets.p, pl.W, pl.Wd, pl.B, pl.logB, ctx->wsum.p); CKL();
    MARK(T_WSUM);
    k_horner<C><<<1, 32, 0, s>>>(ctx->wsum.p, pl.W, pl.Wd, pl.c0, pl.rem, d_out); CKL();
    MARK(T_HORNER);
  }
  if (st) {
    CK(cudaEventRecord(ctx->ev[5], s));
    st->n = n; st->window_bits = pl.c; st->windows = pl.Wd; st->reserved = pl.W; st->buckets_per_window = pl.B; st->tree_rounds = rounds; st->affine_adds = ctx->adds_exact ? ctx->adds_exact : adds; st->affine_adds_round0 = ctx->adds_r0;
  }
  return B200MSM_OK;
}

template <class C> int write_zero(b200msm_ctx* ctx, void* d_out) {
  uint32_t z[3 * C::N]; memset(z, 0, sizeof z);
  for (int i = 0; i < C::N; i++) z[C::N + i] = C::one(i);
  CK(cudaMemcpyAsync(d_out, z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200MSM_OK;
}

int deliver(b200msm_ctx* ctx, const void* d_src, void* user_out, size_t bytes) {
  if (is_device_ptr(user_out)) { CK(cudaMemcpyAsync(user_out, d_src, bytes, cudaMemcpyDefault, ctx->stream)); }      // any device (UVA); stays asynchronous
  else { CK(cudaMemcpyAsync(user_out, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
  return B200MSM_OK;
}

// stage an input: returns a device pointer (the user's if already on device and aligned, else a copy in `buf`)
int stage(b200msm_ctx* ctx, const void* src, size_t bytes, DevBuf& buf, const void** d) {
  if (bytes == 0) { CK(buf.ensure(16)); *d = buf.p; return B200MSM_OK; }
  if (ptr_device(src) == ctx->device && (reinterpret_cast<uintptr_t>(src) & 15) == 0) { *d = src; return B200MSM_OK; }      // another device's memory is copied (UVA)
  CK(buf.ensure(bytes + 16));
  CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyDefault, ctx->stream));
  *d = buf.p; return B200MSM_OK;
}

int msm_entry_impl(b200msm_ctx* ctx, int curve, const void* bases, bool bases_resident, const void* scalars, uint32_t scalar_size, uint64_t n,
                   uint32_t bit0, uint32_t nbits, void* out, b200msm_stats* st, const Precomp* pre) {
  if (!ctx) return B200MSM_E_ARG;
  const void* xs_res = ctx->xs_hint; ctx->xs_hint = nullptr; ctx->cur_xs = nullptr;      // the caller's x-only copy belongs to THIS call only, whatever becomes of it
  if (!curve_ok(curve) || !out || (n && (!bases || !scalars)) || scalar_size == 0) { ctx->err = "bad argument"; return B200MSM_E_ARG; }
  if (n >= (1ull << 31)) { ctx->err = "n must be < 2^31"; return B200MSM_E_UNSUPPORTED; }
  CK(cudaSetDevice(ctx->device));
  const int n8 = n8_of(curve);
  CK(ctx->out.ensure(3 * 96));
  ctx->prof = false;
  if (st) { memset(st, 0, sizeof *st); CK(cudaEventRecord(ctx->ev[0], ctx->stream)); ctx->prof = true; ctx->pused = 0; }
  const uint64_t launches0 = ctx->launches;
  // clip the processed bit range at the scalar end (getChunk's bitsToEnd mask, build_multiexp.js:38-72)
  if (bit0 >= 8 * scalar_size) nbits = 0; else if (bit0 + nbits > 8 * scalar_size) nbits = 8 * scalar_size - bit0;
  if (nbits > 256) { ctx->err = "internal: more than 256 scalar bits per pipeline pass"; return B200MSM_E_UNSUPPORTED; }
  int rc;
  if (n == 0 || nbits == 0) {
    B200_CURVE_SWITCH(curve, rc = write_zero<C>(ctx, ctx->out.p))
    if (rc) return rc;
    return deliver(ctx, ctx->out.p, out, 3 * n8);
  }
  // scalars first (the digit / sort phase needs only them); host-resident bases follow on a separate copy stream and are
  // awaited only where the tree's first round starts gathering points, so their transfer overlaps the sort
  const void* d_sraw; rc = stage(ctx, scalars, (size_t)n * scalar_size, ctx->scalars, &d_sraw); if (rc) return rc;
  const void* d_bases = bases;
  ctx->bases_pending = false;
  if (!bases_resident) {
    const size_t bytes = (size_t)n * 2 * n8;
    if (ptr_device(bases) == ctx->device && (reinterpret_cast<uintptr_t>(bases) & 15) == 0) d_bases = bases;
    else {
      CK(ctx->bases.ensure(bytes + 16));
      if (!ctx->copy_stream) { CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ctx->ev_bases, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming)); }
      CK(cudaEventRecord(ctx->ev_done, ctx->stream));                       // everything issued so far (previous calls may still read ctx->bases)
      CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done, 0));
      CK(cudaMemcpyAsync(ctx->bases.p, bases, bytes, cudaMemcpyDefault, ctx->copy_stream));
      ctx->bases_pending = true; d_bases = ctx->bases.p;
    }
  }
  // x coordinates alone for the forward pass of round 0 (meta_load_x): a resident set brings its copy (built at upload), otherwise it is
  // extracted per call -- behind the transfer on the copy stream when the bases come from the host.  Below 2^19 points the bases fit in L2 as they are.
  if (ctx->opt_xonly && !pre && n >= (1u << 19) && !ctx->prof_noxs) {
    if (xs_res) ctx->cur_xs = xs_res;
    else {
      CK(ctx->xonly.ensure((size_t)n * n8 + 16));
      cudaStream_t xst = ctx->bases_pending ? ctx->copy_stream : ctx->stream;
      B200_CURVE_SWITCH(curve, k_extract_x<C><<<(uint32_t)((n + 255) / 256), 256, 0, xst>>>(d_bases, (uint32_t)n, ctx->xonly.p))
      CKL();
      ctx->cur_xs = ctx->xonly.p;
    }
    ctx->cur_xs_bytes = (uint64_t)n * n8;
  }
  if (ctx->bases_pending) CK(cudaEventRecord(ctx->ev_bases, ctx->copy_stream));
  const uint32_t* d_scal;
  if (scalar_size == 32 && bit0 == 0 && nbits == 256) d_scal = reinterpret_cast<const uint32_t*>(d_sraw);
  else {
    CK(ctx->canon.ensure((size_t)n * 32));
    k_canon_scalars<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint8_t*>(d_sraw), scalar_size, (uint32_t)n, bit0, nbits, ctx->canon.as<uint32_t>()); CKL();
    d_scal = ctx->canon.as<uint32_t>();
  }
  MARK(T_NTAGS);
  if (pre && (pre->nbits != nbits || bit0 != 0)) pre = nullptr;      // the table serves exactly the bit range it was built for
  B200_CURVE_SWITCH(curve, rc = run_pipeline<C>(ctx, d_bases, d_scal, n, nbits, ctx->out.p, st, pre))
  if (rc) { ctx->prof = false; return rc; }
  if (st) CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  rc = deliver(ctx, ctx->out.p, out, 3 * n8); if (rc) return rc;
  if (st) {
    CK(cudaEventRecord(ctx->ev[7], ctx->stream)); CK(cudaEventSynchronize(ctx->ev[7]));
    cudaEventElapsedTime(&st->ms_h2d, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&st->ms_digits_sort, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&st->ms_accumulate, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&st->ms_bucket_reduce, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&st->ms_window_combine, ctx->ev[4], ctx->ev[5]);
    st->ms_host_combine = ctx->opt_combine == 1 ? 0.f : ctx->host_combine_ms;
    cudaEventElapsedTime(&st->ms_d2h, ctx->ev[6], ctx->ev[7]);
    cudaEventElapsedTime(&st->ms_total, ctx->ev[0], ctx->ev[7]);
    uint32_t total = 0;
    uint32_t nb = st->reserved * st->buckets_per_window;
    CK(cudaMemcpy(&total, ctx->offsets.as<uint32_t>() + nb, 4, cudaMemcpyDeviceToHost));
    st->pairs = total;
    float acc[T_NTAGS + 1] = {0};
    for (size_t i = 1; i < ctx->pused; i++) { float ms = 0; cudaEventElapsedTime(&ms, ctx->pev[i - 1], ctx->pev[i]); acc[ctx->ptag[i]] += ms; }
    st->ms_k_sort = acc[T_SORT]; st->ms_k_plan = acc[T_PLAN]; st->ms_k_tree_fwd = acc[T_TREE_FWD]; st->ms_k_inv_tree = acc[T_INV_TREE];
    st->ms_k_tree_bwd = acc[T_TREE_BWD] + acc[T_TREE_BWD0]; st->ms_k_tree_bwd_round0 = acc[T_TREE_BWD0]; st->ms_k_finish = acc[T_FINISH]; st->ms_k_fold = acc[T_FOLD]; st->ms_k_wsum = acc[T_WSUM]; st->ms_k_horner = acc[T_HORNER];
    st->launches = ctx->launches - launches0;
    ctx->prof = false;
  }
  return B200MSM_OK;
}

// acc = 2^shift * acc + part on Jacobian Montgomery points (device buffers of 3*n8 bytes): the Horner step between two 256-bit scalar slices
template <class C>
__global__ void k_shift_add_jacobian(void* __restrict__ acc_jac, const void* __restrict__ part_jac, uint32_t shift) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  XYZZ<C> a, p;
  { const char* s = reinterpret_cast<const char*>(acc_jac); Fe<C::N> X, Y, Z; fe_load_cg<C>(X, s); fe_load_cg<C>(Y, s + 4 * C::N); fe_load_cg<C>(Z, s + 8 * C::N);
    if (fe_is_zero<C>(Z)) xyzz_set_inf<C>(a); else { a.x = X; a.y = Y; fe_sqr<C>(a.zz, Z); fe_mul<C>(a.zzz, a.zz, Z); } }
  { const char* s = reinterpret_cast<const char*>(part_jac); Fe<C::N> X, Y, Z; fe_load_cg<C>(X, s); fe_load_cg<C>(Y, s + 4 * C::N); fe_load_cg<C>(Z, s + 8 * C::N);
    if (fe_is_zero<C>(Z)) xyzz_set_inf<C>(p); else { p.x = X; p.y = Y; fe_sqr<C>(p.zz, Z); fe_mul<C>(p.zzz, p.zz, Z); } }
  for (uint32_t k = 0; k < shift; k++) { XYZZ<C> d; xyzz_dbl<C>(d, a); a = d; }
  xyzz_add<C>(a, p);
  Fe<C::N> X, Y, Z; xyzz_to_jacobian<C>(X, Y, Z, a);
  char* o = reinterpret_cast<char*>(acc_jac);
  fe_store<C>(o, X); fe_store<C>(o + 4 * C::N, Y); fe_store<C>(o + 8 * C::N, Z);
}

// One MSM over scalar bits [bit0, bit0 + nbits).  Bit ranges wider than 256 bits (scalar_size > 32: the reference's g1m_multiexpAffine takes any
// scalarSize, build_multiexp.js:251-371) run as 256-bit slices from the top down, combined by Horner: result = 2^256 * result + slice.
int msm_entry(b200msm_ctx* ctx, int curve, const void* bases, bool bases_resident, const void* scalars, uint32_t scalar_size, uint64_t n,
              uint32_t bit0, uint32_t nbits, void* out, b200msm_stats* st, const Precomp* pre = nullptr) {
  if (!ctx) return B200MSM_E_ARG;
  uint32_t eff = nbits;
  if (scalar_size && bit0 < 8ull * scalar_size && (uint64_t)bit0 + nbits > 8ull * scalar_size) eff = 8 * scalar_size - bit0;
  int rc;
  if (eff <= 256 || !curve_ok(curve) || !out || n == 0 || !bases || !scalars) rc = msm_entry_impl(ctx, curve, bases, bases_resident, scalars, scalar_size, n, bit0, nbits, out, st, pre);
  else {
    // keep the bases on the device across the slices
    cudaError_t ce = cudaSetDevice(ctx->device);
    const int n8 = n8_of(curve);
    const void* d_bases = bases;
    rc = B200MSM_OK;
    if (ce != cudaSuccess) { ctx->err = "cudaSetDevice failed"; rc = B200MSM_E_CUDA; }
    DevBuf keep_bases, acc, part;
    if (!rc && !bases_resident && !(ptr_device(bases) == ctx->device && (reinterpret_cast<uintptr_t>(bases) & 15) == 0)) {
      ce = keep_bases.ensure((size_t)n * 2 * n8 + 16);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(keep_bases.p, bases, (size_t)n * 2 * n8, cudaMemcpyDefault, ctx->stream);
      if (ce != cudaSuccess) { ctx->err = std::string("staging bases: ") + cudaGetErrorString(ce); rc = ce == cudaErrorMemoryAllocation ? B200MSM_E_NOMEM : B200MSM_E_CUDA; }
      d_bases = keep_bases.p;
    }
    if (!rc && (acc.ensure(3 * 96) != cudaSuccess || part.ensure(3 * 96) != cudaSuccess)) { ctx->err = "out of device memory"; rc = B200MSM_E_NOMEM; }
    const uint32_t slices = (eff + 255) / 256;
    for (int k = (int)slices - 1; k >= 0 && !rc; k--) {
      const uint32_t b0 = bit0 + 256u * (uint32_t)k, nb = std::min<uint32_t>(256u, eff - 256u * (uint32_t)k);
      rc = msm_entry_impl(ctx, curve, d_bases, true, scalars, scalar_size, n, b0, nb, k == (int)slices - 1 ? acc.p : part.p, nullptr, nullptr);
      if (!rc && k != (int)slices - 1) {
        B200_CURVE_SWITCH(curve, k_shift_add_jacobian<C><<<1, 32, 0, ctx->stream>>>(acc.p, part.p, 256u))
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { ctx->err = "k_shift_add_jacobian launch failed"; rc = B200MSM_E_CUDA; }
      }
    }
    if (!rc) rc = deliver(ctx, acc.p, out, 3 * (size_t)n8);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "stream synchronize failed"; rc = B200MSM_E_CUDA; }
    cudaStreamSynchronize(ctx->stream);
    keep_bases.release(); acc.release(); part.release();
    if (st) memset(st, 0, sizeof *st);
  }
  ctx->prof = false;                     // the phase profiler never stays armed after an error return
  return rc;
}

// window table for resident bases: rows w = 1..Wd-1 hold 2^(bit offset of window w) * P_i (see k_table_double)
template <class C>
int build_window_table(b200msm_ctx* ctx, void* table, uint64_t n, uint32_t c0, uint32_t rem, uint32_t Wd) {
  const size_t pt = 8 * C::N;
  CK(ctx->acc_a.ensure(n * 16 * C::N + 16));
  const uint32_t g = (uint32_t)((n + 127) / 128);
  constexpr int GROUP = 16;
  const uint32_t g2 = (uint32_t)(((n + GROUP - 1) / GROUP + 127) / 128);
  k_table_init<C><<<g, 128, 0, ctx->stream>>>(table, (uint32_t)n, ctx->acc_a.p); CKL();
  for (uint32_t w = 1; w < Wd; w++) {
    const uint32_t cw = c0 + ((w - 1) < rem ? 1u : 0u);
    k_table_double<C><<<g, 128, 0, ctx->stream>>>(ctx->acc_a.p, (uint32_t)n, cw); CKL();
    k_xyzz_to_affine<C, GROUP><<<g2, 128, 0, ctx->stream>>>(ctx->acc_a.p, (uint32_t)n, reinterpret_cast<char*>(table) + (size_t)w * n * pt); CKL();
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return B200MSM_OK;
}

}  // namespace

// =================================================================== internal interface for the library's other translation units (internal.h)
cudaStream_t b200msm_internal_stream(b200msm_ctx* ctx) { return ctx->stream; }
int b200msm_internal_device(b200msm_ctx* ctx) { return ctx->device; }
void b200msm_internal_count_launches(b200msm_ctx* ctx, uint64_t k) { ctx->launches += k; }
void b200msm_internal_set_error(b200msm_ctx* ctx, const char* msg) { ctx->err = msg ? msg : ""; }

// =================================================================== C ABI
extern "C" {

const char* b200msm_version(void) { return "b200msm 0.1 (sm_100a)"; }

const char* b200msm_strerror(int s) {
  switch (s) {
    case B200MSM_OK: return "ok";
    case B200MSM_E_ARG: return "invalid argument";
    case B200MSM_E_CUDA: return "CUDA error (no usable GPU, or a kernel/driver failure)";
    case B200MSM_E_NOMEM: return "out of device memory";
    case B200MSM_E_UNSUPPORTED: return "unsupported size or parameter";
    default: return "unknown status";
  }
}
const char* b200msm_last_error(const b200msm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b200msm_create(b200msm_ctx** out, int device_id) {
  if (!out) return B200MSM_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return B200MSM_E_CUDA; }
  int dev = device_id;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return B200MSM_E_CUDA; }
  if (dev >= ndev) return B200MSM_E_ARG;
  if (cudaSetDevice(dev) != cudaSuccess) return B200MSM_E_CUDA;
  b200msm_ctx* ctx = new b200msm_ctx();
  ctx->device = dev;
  { size_t fr = 0, tot = 0; if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) ctx->total_mem = tot; else ctx->total_mem = (size_t)64 << 30; }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return B200MSM_E_CUDA; }
  ctx->own_stream = true;
  for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return B200MSM_E_CUDA; }
  if (cudaMallocHost(&ctx->h_pinned, 2048 * sizeof(uint32_t)) != cudaSuccess) { delete ctx; return B200MSM_E_CUDA; }
  if (cudaEventCreateWithFlags(&ctx->ev_plan, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_sorted, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return B200MSM_E_CUDA; }
  *out = ctx;
  return B200MSM_OK;
}

void b200msm_destroy(b200msm_ctx* ctx) {
  if (!ctx) return;
  for (size_t g = 1; g < ctx->devs.size(); g++) b200msm_destroy(ctx->devs[g]);      // a multi context owns the contexts of its other devices
  ctx->devs.clear(); ctx->mres.clear();
  b200ntt_release(ctx);
  for (b200msm_ctx* w : ctx->workers) b200msm_destroy(w);
  ctx->workers.clear();
  delete ctx->pool; ctx->pool = nullptr;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (DevBuf* b : {&ctx->bases, &ctx->scalars, &ctx->canon, &ctx->counts, &ctx->offsets, &ctx->ranks, &ctx->tiles, &ctx->sorted, &ctx->buckets,
                    &ctx->wsum, &ctx->out, &ctx->misc, &ctx->acc_a, &ctx->acc_b, &ctx->acc_c, &ctx->acc_d, &ctx->acc_e, &ctx->jac_in, &ctx->jac_affine, &ctx->xonly}) b->release();
  for (auto& ln : ctx->lane) {
    for (DevBuf* b : {&ln.offs, &ln.tiles, &ln.bid, &ln.pa, &ln.pb, &ln.prefix, &ln.prod, &ln.lvlprefix, &ln.others, &ln.meta, &ln.carry, &ln.btree}) b->release();
    if (ln.done) cudaEventDestroy(ln.done);
    if (ln.stream) cudaStreamDestroy(ln.stream);
  }
  if (ctx->ev_plan) cudaEventDestroy(ctx->ev_plan);
  if (ctx->ev_sorted) cudaEventDestroy(ctx->ev_sorted);
  if (ctx->ev_bases) cudaEventDestroy(ctx->ev_bases);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& e : ctx->gev) cudaEventDestroy(e);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_folded) cudaFreeHost(ctx->h_folded);
  for (auto& kv : ctx->residents) { cudaFree(kv.second.d); if (kv.second.dx) cudaFree(kv.second.dx); }
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->pev) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int b200msm_set_stream(b200msm_ctx* ctx, void* cuda_stream) {
  if (!ctx) return B200MSM_E_ARG;
  if (ctx->devs.size() > 1) { ctx->err = "a multi-device context owns its streams (one per device)"; return B200MSM_E_UNSUPPORTED; }
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);       // work issued on the old stream (cached NTT twiddle tables, resident uploads) completes before the switch
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream); ctx->own_stream = false;
  return B200MSM_OK;
}
int b200msm_synchronize(b200msm_ctx* ctx) {
  if (!ctx) return B200MSM_E_ARG;
  for (size_t g = 1; g < ctx->devs.size(); g++) { CK(cudaSetDevice(ctx->devs[g]->device)); CK(cudaStreamSynchronize(ctx->devs[g]->stream)); }
  CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream)); return B200MSM_OK;
}

int b200msm_set_option(b200msm_ctx* ctx, const char* key, int64_t v) {
  if (!ctx || !key) return B200MSM_E_ARG;
  for (size_t g = 1; g < ctx->devs.size(); g++) { int rc = b200msm_set_option(ctx->devs[g], key, v); if (rc) return rc; }
  if (!strcmp(key, "multi_min_points")) { if (v < 1) return B200MSM_E_ARG; ctx->opt_multi_min = v; return B200MSM_OK; }
  if (!strcmp(key, "multi_replicate")) { ctx->opt_multi_replicate = v != 0; return B200MSM_OK; }
  if (!strcmp(key, "window_bits")) { if (v < 0 || v > 24) return B200MSM_E_ARG; ctx->opt_window_bits = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "accumulate")) { if (v < 0 || v > 2) return B200MSM_E_ARG; ctx->opt_accumulate = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "tree_rounds")) { ctx->opt_tree_rounds = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "group_pairs")) { if (v < 0) return B200MSM_E_ARG; ctx->opt_group_pairs = v; return B200MSM_OK; }
  if (!strcmp(key, "persist")) { if (v < 0) return B200MSM_E_ARG; ctx->opt_persist = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "ba_k")) { if (v < 0 || v > 64) return B200MSM_E_ARG; ctx->opt_ba_k = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "pt_k")) { if (v < 2 || v > 64) return B200MSM_E_ARG; ctx->opt_pt_k = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "probe_sqr")) { ctx->probe_sqr = v != 0; return B200MSM_OK; }
#if defined(B200_EXPERIMENTS)
  if (!strcmp(key, "probe29")) { ctx->probe29 = v != 0; return B200MSM_OK; }
#endif
  if (!strcmp(key, "meta_upfront")) { ctx->opt_meta_upfront = v != 0; return B200MSM_OK; }
  if (!strcmp(key, "xonly")) { ctx->opt_xonly = v != 0; return B200MSM_OK; }
  if (!strcmp(key, "group_plan")) { if (v < 0) return B200MSM_E_ARG; ctx->opt_group_plan = v; return B200MSM_OK; }
  if (!strcmp(key, "block_tree")) { ctx->opt_block_tree = v != 0; return B200MSM_OK; }
  if (!strcmp(key, "fused_round")) { if (v < 0 || v > 3) return B200MSM_E_ARG; ctx->opt_fused = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "fused_grid")) { if (v < 1 || v > 65536) return B200MSM_E_ARG; ctx->opt_fused_grid = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "fused_tiles")) { if (v < 1 || v > 65536) return B200MSM_E_ARG; ctx->opt_fused_tiles = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "fused_kmax")) { if (v < 1 || v > 64) return B200MSM_E_ARG; ctx->opt_fused_kmax = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "persist_fwd")) { if (v < 0 || v > 4096) return B200MSM_E_ARG; ctx->opt_persist_fwd = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "groups")) { if (v < 0 || v > 64) return B200MSM_E_ARG; ctx->opt_groups = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "group_small")) { if (v < 5 || v > 100) return B200MSM_E_ARG; ctx->opt_group_small = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "fold_cluster")) { if (v < 0 || v > 2) return B200MSM_E_ARG; ctx->opt_fold_cluster = (int)v; return B200MSM_OK; }      // 0 = one CTA per live block, 1 = a cluster per live block, 2 = the same with four lanes per addition
  if (!strcmp(key, "issue_threads")) { ctx->opt_issue_threads = v != 0; return B200MSM_OK; }
  if (!strcmp(key, "sort_groups")) { ctx->opt_sort_groups = v != 0; return B200MSM_OK; }
  if (!strcmp(key, "lanes")) { if (v < 1 || v > MAX_LANES) return B200MSM_E_ARG; ctx->opt_lanes = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "probe_smem")) { if (v < 0 || v > 200 * 1024) return B200MSM_E_ARG; ctx->opt_probe_smem = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "batch_blocking")) { ctx->opt_batch_blocking = v != 0; return B200MSM_OK; }      // takes effect for workers created afterwards
  if (!strcmp(key, "batch_lanes")) { if (v < 1 || v > MAX_LANES) return B200MSM_E_ARG; ctx->opt_batch_lanes = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "batch_workers")) { if (v < 1 || v > 16) return B200MSM_E_ARG; ctx->opt_batch_workers = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "subslots")) { if (v < 0 || v > 256 || (v & (v - 1))) return B200MSM_E_ARG; ctx->opt_subslots = (int)v; return B200MSM_OK; }
  if (!strcmp(key, "combine")) { if (v < 0 || v > 1) return B200MSM_E_ARG; ctx->opt_combine = (int)v; return B200MSM_OK; }
  return B200MSM_E_ARG;
}

// =================================================================== single-device cores of the MSM entry points
// (the public functions below dispatch: a context made by b200msm_create_multi shards the same calls over its devices)

// Jacobian bases (n8b = 3*n8, build_curve_jacobian_a0.js:1429): one batched conversion to affine on the device, then the affine pipeline
static int msm_jacobian(b200msm_ctx* ctx, int curve, const void* bases_jac, const void* scalars, uint32_t scalar_size, uint64_t n,
                        uint32_t bit0, uint32_t nbits, void* out) {
  if (!ctx) return B200MSM_E_ARG;
  if (!curve_ok(curve) || !out || (n && (!bases_jac || !scalars)) || scalar_size == 0 || n >= (1ull << 31)) { ctx->err = "bad argument"; return B200MSM_E_ARG; }
  if (n == 0) return msm_entry(ctx, curve, bases_jac, false, scalars, scalar_size, 0, bit0, nbits, out, nullptr);
  CK(cudaSetDevice(ctx->device));
  const size_t n8 = n8_of(curve);
  const void* d_in; int rc = stage(ctx, bases_jac, n * 3 * n8, ctx->jac_in, &d_in); if (rc) return rc;
  CK(ctx->jac_affine.ensure(n * 2 * n8 + 16));
  constexpr int GROUP = 8;
  const uint32_t g2 = (uint32_t)(((n + GROUP - 1) / GROUP + 127) / 128);
  B200_CURVE_SWITCH(curve, k_jacobian_to_affine<C, GROUP><<<g2, 128, 0, ctx->stream>>>((const uint8_t*)d_in, (uint32_t)n, ctx->jac_affine.as<uint8_t>()))
  CKL();
  return msm_entry(ctx, curve, ctx->jac_affine.p, false, scalars, scalar_size, n, bit0, nbits, out, nullptr);
}

static int single_upload(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, uint64_t* handle) {
  if (!ctx || !handle || !curve_ok(curve) || (n && !bases)) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  size_t bytes = (size_t)n * 2 * n8_of(curve);
  void* d = nullptr;
  CK(cudaMalloc(&d, bytes + 16));
  cudaError_t e = cudaMemcpyAsync(d, bases, bytes, cudaMemcpyDefault, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { cudaFree(d); ctx->err = cudaGetErrorString(e); return B200MSM_E_CUDA; }
  void* dx = nullptr;
  if (n >= (1u << 19) && n < (1ull << 31)) {      // x coordinates alone for the forward pass of round 0 (a failed allocation just leaves the per-call copy in charge)
    if (cudaMalloc(&dx, (size_t)n * n8_of(curve) + 16) == cudaSuccess) {
      B200_CURVE_SWITCH(curve, k_extract_x<C><<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(d, (uint32_t)n, dx))
      ctx->launches++;
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(dx); dx = nullptr; }
    } else { cudaGetLastError(); dx = nullptr; }
  }
  uint64_t h = ctx->next_handle++;
  Resident r{curve, n, d}; r.dx = dx;
  ctx->residents[h] = r;
  *handle = h;
  return B200MSM_OK;
}
static int single_upload_windowed(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, uint32_t scalar_size, uint32_t window_bits, uint64_t* handle) {
  if (!ctx || !handle || !curve_ok(curve) || !n || !bases || scalar_size == 0 || scalar_size > 32 || window_bits > 24) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const uint32_t nbits = 8 * scalar_size;
  uint32_t ct = window_bits;
  if (ct == 0) { uint32_t lg = 0; while ((2ull << lg) <= n) lg++; ct = lg <= 11 ? 15 : lg <= 18 ? 18 : std::min<uint32_t>(lg, 20); }      // one bucket array: wider windows than the per-window plan; small sets want ~4 points per bucket so that no tree round is needed (measured, ms at auto = log2 n -> now: 2^12 1.19 -> 0.68, 2^14 1.35 -> 0.83, 2^16 1.51 -> 1.35; 2^18 -> 18, 2^20 -> 20 as before)
  if (ct < 2) ct = 2;
  if (ct > nbits) ct = nbits;
  const uint32_t Wd = (nbits + ct - 1) / ct, c0 = nbits / Wd, rem = nbits - c0 * Wd;
  if (n * Wd >= (1ull << 31)) { ctx->err = "window table too large for 31-bit point indices"; return B200MSM_E_UNSUPPORTED; }
  const size_t pt = 2 * (size_t)n8_of(curve);
  void* d = nullptr;
  { cudaError_t e = cudaMalloc(&d, (size_t)n * Wd * pt + 16);
    if (e != cudaSuccess) { cudaGetLastError(); ctx->err = "window table does not fit in device memory"; return B200MSM_E_NOMEM; } }
  cudaError_t e = cudaMemcpyAsync(d, bases, (size_t)n * pt, cudaMemcpyDefault, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { cudaFree(d); ctx->err = cudaGetErrorString(e); return B200MSM_E_CUDA; }
  int rc; B200_CURVE_SWITCH(curve, rc = build_window_table<C>(ctx, d, n, c0, rem, Wd))
  if (rc) { cudaFree(d); return rc; }
  uint64_t h = ctx->next_handle++;
  Resident r{curve, n, d}; r.t_nbits = nbits; r.t_c0 = c0; r.t_rem = rem; r.t_Wd = Wd;
  ctx->residents[h] = r;
  *handle = h;
  return B200MSM_OK;
}
static int single_free(b200msm_ctx* ctx, uint64_t handle) {
  auto it = ctx->residents.find(handle);
  if (it == ctx->residents.end()) return B200MSM_E_ARG;
  cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
  for (b200msm_ctx* w : ctx->workers) cudaStreamSynchronize(w->stream);
  cudaFree(it->second.d); if (it->second.dx) cudaFree(it->second.dx); ctx->residents.erase(it);
  return B200MSM_OK;
}
// MSM over the resident points [first, first + n) of a handle (first > 0: one device's share of a replicated base set)
static int single_resident(b200msm_ctx* ctx, uint64_t handle, uint64_t first, const void* scalars, uint32_t scalar_size, uint64_t n, void* out, b200msm_stats* stats) {
  auto it = ctx->residents.find(handle);
  if (it == ctx->residents.end() || first + n > it->second.n) { ctx->err = "unknown handle or n larger than the uploaded base count"; return B200MSM_E_ARG; }
  const Resident& r = it->second;
  Precomp pre{(uint32_t)r.n, r.t_nbits, r.t_c0, r.t_rem, r.t_Wd};          // a table row is r.n points long whatever the offset
  const char* d = reinterpret_cast<const char*>(r.d) + first * 2 * (size_t)n8_of(r.curve);
  ctx->xs_hint = r.dx ? reinterpret_cast<const char*>(r.dx) + first * (size_t)n8_of(r.curve) : nullptr;
  return msm_entry(ctx, r.curve, d, true, scalars, scalar_size, n, 0, 8 * scalar_size, out, stats, r.t_Wd ? &pre : nullptr);
}

// MSMs j = first, first + stride, ... < count of a batch over the same resident bases (BASELINE config 5: streams of equal-size MSMs).  They are
// spread over `batch_workers` sub-contexts -- each with its own stream and scratch, driven by its own host thread -- so that the latency-bound
// parts of one MSM (sort read-back, inversion tails, host window combination) overlap the throughput-bound kernels of the others.
static int single_batch(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size, uint64_t n, uint32_t count, uint32_t first, uint32_t stride, void* out) {
  auto it = ctx->residents.find(handle);
  if (it == ctx->residents.end() || n > it->second.n || !out || scalar_size == 0 || (n && !scalars)) { ctx->err = "bad handle or argument"; return B200MSM_E_ARG; }
  if (first >= count) return B200MSM_OK;
  const Resident r = it->second;
  const int n8 = n8_of(r.curve);
  const uint32_t mine = (count - first + stride - 1) / stride;
  const uint32_t K = std::min<uint32_t>((uint32_t)ctx->opt_batch_workers, mine);
  while (ctx->workers.size() < K) {
    b200msm_ctx* w = nullptr; int rc = b200msm_create(&w, ctx->device);
    if (rc) { ctx->err = "cannot create a batch worker context"; return rc; }
    if (ctx->opt_batch_blocking) {      // host waits sleep instead of spinning: for hosts where the workers of several processes outnumber the cores (measured: one process, 16 cores: 4 % slower)
      w->blocking_waits = true;
      cudaEventDestroy(w->ev_plan); w->ev_plan = nullptr;
      if (cudaEventCreateWithFlags(&w->ev_plan, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) { b200msm_destroy(w); ctx->err = "cannot create a batch worker context"; return B200MSM_E_CUDA; }
    }
    ctx->workers.push_back(w);
  }
  for (uint32_t k = 0; k < K; k++) { copy_options(ctx->workers[k], ctx); ctx->workers[k]->mem_share = 1.0 / K; }      // workers inherit the tuning options of the parent and share its memory budget
  CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream));       // inputs produced on the caller's stream are complete
  Precomp pre{(uint32_t)r.n, r.t_nbits, r.t_c0, r.t_rem, r.t_Wd};
  const uint32_t nbits = 8 * scalar_size;
  std::vector<int> rcs(K, 0);
  std::vector<std::thread> th;
  for (uint32_t k = 0; k < K; k++) th.emplace_back([&, k]() {
    b200msm_ctx* w = ctx->workers[k];
    for (uint32_t t = k; t < mine; t += K) {
      const uint64_t j = first + (uint64_t)t * stride;
      const char* sc = reinterpret_cast<const char*>(scalars) + (size_t)j * n * scalar_size;
      char* o = reinterpret_cast<char*>(out) + (size_t)j * 3 * n8;
      w->xs_hint = r.dx;
      int rc = msm_entry(w, r.curve, r.d, true, sc, scalar_size, n, 0, nbits, o, nullptr, r.t_Wd ? &pre : nullptr);
      if (rc) { rcs[k] = rc; return; }
    }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) rcs[k] = B200MSM_E_CUDA;
  });
  for (auto& t : th) t.join();
  int first_rc = 0; std::string msg;                                    // every failing worker is reported, not only the first
  for (uint32_t k = 0; k < K; k++) if (rcs[k]) { if (!first_rc) first_rc = rcs[k]; msg += (msg.empty() ? "" : "; ") + ("batch worker " + std::to_string(k) + ": " + ctx->workers[k]->err); }
  if (first_rc) { ctx->err = msg; return first_rc; }
  return B200MSM_OK;
}

// =================================================================== multi-device contexts
// b200msm_create_multi makes ONE context that owns G single-device contexts (device g: its own streams, scratch and host thread).  An MSM is a sum
// over independent (point, scalar) pairs, so the calls shard by POINT RANGE exactly as ffjavascript shards g1m_multiexpAffine over its workers
// (slices of the inputs per worker, partial results added: wasmcurves/src/build_multiexp.js:319-369 is the per-worker part): device g pulls its
// slice of the caller's host buffers over its own PCIe link, runs the whole single-GPU pipeline and returns ONE partial G1 point (3*n8 bytes);
// the G partials are added on the host (host_ec.h, ~1 us each).  No bases or scalars ever cross NVLink, so there is no collective on the data path.
}  // extern "C"  (the templates below need C++ linkage)
namespace {

void shard_of(uint64_t n, uint32_t g, uint32_t G, uint64_t* lo, uint64_t* cnt) {      // contiguous, balanced: the first n % G devices get one extra point
  const uint64_t base = n / G, extra = n % G;
  *lo = g * base + std::min<uint64_t>(g, extra); *cnt = base + (g < extra ? 1 : 0);
}

// f(g) -> status, run for g = 0..G-1 concurrently (g = 0 on the calling thread); every failing device is reported
template <class F> int multi_run(b200msm_ctx* ctx, uint32_t G, F&& f) {
  std::vector<int> rcs(G, 0);
  std::vector<std::thread> th;
  for (uint32_t g = 1; g < G; g++) th.emplace_back([&, g]() { rcs[g] = f(g); });
  rcs[0] = f(0);
  for (auto& t : th) t.join();
  int first_rc = 0; std::string msg;
  for (uint32_t g = 0; g < G; g++) if (rcs[g]) { if (!first_rc) first_rc = rcs[g];
    msg += (msg.empty() ? "" : "; ") + ("device " + std::to_string(ctx->devs[g]->device) + ": " + (g ? ctx->devs[g]->err : ctx->err)); }
  if (first_rc) ctx->err = msg;
  cudaSetDevice(ctx->device);
  return first_rc;
}

// out (3*n8 bytes, host) = sum of G Jacobian Montgomery points (g1m_add chain, build_curve_jacobian_a0.js:541-658) on the host
template <class C> void host_sum_jacobian(const uint8_t* parts, uint32_t G, uint8_t* out) {
  using HF = typename HostField<C>::type; const HF f = HostField<C>::make();
  constexpr int W = HF::W;
  b200host::XYZZ<W> acc; b200host::set_inf(f, acc);
  for (uint32_t g = 0; g < G; g++) {
    b200host::Fe<W> X, Y, Z;
    memcpy(X.l, parts + (size_t)g * 24 * W, 8 * W); memcpy(Y.l, parts + (size_t)g * 24 * W + 8 * W, 8 * W); memcpy(Z.l, parts + (size_t)g * 24 * W + 16 * W, 8 * W);
    if (b200host::is_zero<W>(Z)) continue;
    b200host::XYZZ<W> p; p.x = X; p.y = Y; b200host::sqr(f, p.zz, Z); b200host::mul(f, p.zzz, p.zz, Z);
    b200host::padd(f, acc, p);
  }
  b200host::Combiner<HF> cb; cb.f = f; cb.acc = acc; cb.cur = 0;
  cb.finish(reinterpret_cast<uint64_t*>(out));
}

int multi_deliver(b200msm_ctx* ctx, int curve, const std::vector<uint8_t>& parts, uint32_t G, void* out) {
  const size_t sz = 3 * (size_t)n8_of(curve);
  std::vector<uint64_t> res(sz / 8);
  B200_CURVE_SWITCH(curve, host_sum_jacobian<C>(parts.data(), G, reinterpret_cast<uint8_t*>(res.data())))
  if (is_device_ptr(out)) { CK(cudaSetDevice(ctx->device)); CK(cudaMemcpy(out, res.data(), sz, cudaMemcpyDefault)); }
  else memcpy(out, res.data(), sz);
  return B200MSM_OK;
}

// devices that take part in an MSM of n points: small problems stay on fewer devices (sharding a latency-bound MSM only adds the slowest tail)
uint32_t multi_width(const b200msm_ctx* ctx, uint64_t n) {
  const uint64_t G = ctx->devs.size(), m = (uint64_t)std::max<int64_t>(1, ctx->opt_multi_min);
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(G, n / m));
}

int multi_msm(b200msm_ctx* ctx, int curve, const void* bases, bool jacobian, const void* scalars, uint32_t scalar_size, uint64_t n, uint32_t bit0, uint32_t nbits, void* out) {
  if (!curve_ok(curve) || !out || (n && (!bases || !scalars)) || scalar_size == 0) { ctx->err = "bad argument"; return B200MSM_E_ARG; }
  const uint32_t G = multi_width(ctx, n);
  if (G <= 1) return jacobian ? msm_jacobian(ctx, curve, bases, scalars, scalar_size, n, bit0, nbits, out) : msm_entry(ctx, curve, bases, false, scalars, scalar_size, n, bit0, nbits, out, nullptr);
  const size_t n8 = n8_of(curve), pt = (jacobian ? 3 : 2) * n8;
  std::vector<uint8_t> parts((size_t)G * 3 * n8);
  int rc = multi_run(ctx, G, [&](uint32_t g) {
    uint64_t lo, cnt; shard_of(n, g, G, &lo, &cnt);
    const char* b = reinterpret_cast<const char*>(bases) + lo * pt; const char* sc = reinterpret_cast<const char*>(scalars) + lo * scalar_size;
    uint8_t* o = parts.data() + (size_t)g * 3 * n8;
    return jacobian ? msm_jacobian(ctx->devs[g], curve, b, sc, scalar_size, cnt, bit0, nbits, o) : msm_entry(ctx->devs[g], curve, b, false, sc, scalar_size, cnt, bit0, nbits, o, nullptr);
  });
  if (rc) return rc;
  return multi_deliver(ctx, curve, parts, G, out);
}

int multi_upload(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, bool windowed, uint32_t scalar_size, uint32_t window_bits, uint64_t* handle) {
  if (!handle || !curve_ok(curve) || (n && !bases)) return B200MSM_E_ARG;
  const uint32_t G = (uint32_t)ctx->devs.size();
  MultiResident mr; mr.curve = curve; mr.n = n; mr.replicated = ctx->opt_multi_replicate != 0; mr.h.assign(G, 0); mr.lo.assign(G, 0); mr.cnt.assign(G, 0);
  const size_t pt = 2 * (size_t)n8_of(curve);
  for (uint32_t g = 0; g < G; g++) { if (mr.replicated) { mr.lo[g] = 0; mr.cnt[g] = n; } else shard_of(n, g, G, &mr.lo[g], &mr.cnt[g]); }
  int rc = multi_run(ctx, G, [&](uint32_t g) {
    if (mr.cnt[g] == 0) return (int)B200MSM_OK;
    const char* b = reinterpret_cast<const char*>(bases) + mr.lo[g] * pt;
    return windowed ? single_upload_windowed(ctx->devs[g], curve, b, mr.cnt[g], scalar_size, window_bits, &mr.h[g]) : single_upload(ctx->devs[g], curve, b, mr.cnt[g], &mr.h[g]);
  });
  if (rc) { for (uint32_t g = 0; g < G; g++) if (mr.h[g]) single_free(ctx->devs[g], mr.h[g]); return rc; }
  const uint64_t h = ctx->next_handle++;
  ctx->mres[h] = mr; *handle = h;
  return B200MSM_OK;
}

int multi_resident(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size, uint64_t n, void* out, b200msm_stats* stats) {
  auto it = ctx->mres.find(handle);
  if (it == ctx->mres.end() || n > it->second.n || !out || scalar_size == 0 || (n && !scalars)) { ctx->err = "unknown handle, or n larger than the uploaded base count"; return B200MSM_E_ARG; }
  const MultiResident& mr = it->second;
  const uint32_t Gall = (uint32_t)ctx->devs.size(), G = mr.replicated ? multi_width(ctx, n) : Gall;
  const size_t n8 = n8_of(mr.curve);
  std::vector<uint8_t> parts((size_t)G * 3 * n8);
  int rc = multi_run(ctx, G, [&](uint32_t g) {
    uint64_t lo, cnt, first;
    if (mr.replicated) { shard_of(n, g, G, &lo, &cnt); first = lo; }                    // every device holds all points: balanced shares of the first n
    else { lo = mr.lo[g]; cnt = n > lo ? std::min<uint64_t>(mr.cnt[g], n - lo) : 0; first = 0; }
    uint8_t* o = parts.data() + (size_t)g * 3 * n8;
    if (cnt == 0 || !mr.h[g]) {          // nothing of this MSM lives on device g: its partial is the canonical zero (0, R mod q, 0)
      uint32_t one[24] = {0};
      B200_CURVE_SWITCH(mr.curve, for (int i = 0; i < C::N; i++) one[i] = C::one(i))
      memset(o, 0, 3 * n8); memcpy(o + n8, one, n8); return (int)B200MSM_OK; }
    const char* sc = reinterpret_cast<const char*>(scalars) + lo * scalar_size;
    return single_resident(ctx->devs[g], mr.h[g], first, sc, scalar_size, cnt, o, g == 0 ? stats : nullptr);
  });
  if (rc) return rc;
  return multi_deliver(ctx, mr.curve, parts, G, out);
}

int multi_batch(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size, uint64_t n, uint32_t count, void* out) {
  auto it = ctx->mres.find(handle);
  if (it == ctx->mres.end() || n > it->second.n || !out || scalar_size == 0 || (n && !scalars)) { ctx->err = "bad handle or argument"; return B200MSM_E_ARG; }
  const MultiResident mr = it->second;
  if (!mr.replicated) {      // sharded bases: every MSM of the batch runs across all devices, one after the other
    const size_t n8 = n8_of(mr.curve);
    for (uint32_t j = 0; j < count; j++) {
      int rc = multi_resident(ctx, handle, reinterpret_cast<const char*>(scalars) + (size_t)j * n * scalar_size, scalar_size, n, reinterpret_cast<char*>(out) + (size_t)j * 3 * n8, nullptr);
      if (rc) return rc;
    }
    return B200MSM_OK;
  }
  const uint32_t G = std::min<uint32_t>((uint32_t)ctx->devs.size(), std::max<uint32_t>(1, count));      // replicated bases: MSM j on device j mod G (replicas, no exchange)
  return multi_run(ctx, G, [&](uint32_t g) { return single_batch(ctx->devs[g], mr.h[g], scalars, scalar_size, n, count, g, G, out); });
}

}  // namespace

extern "C" {

int b200msm_create_multi(b200msm_ctx** out, const int* device_ids, int n_devices) {
  if (!out || !device_ids || n_devices < 1 || n_devices > 64) return B200MSM_E_ARG;
  *out = nullptr;
  for (int i = 0; i < n_devices; i++) if (device_ids[i] < 0) return B200MSM_E_ARG;      // an ordinal may repeat: every entry gets its own context (two shards on one GPU, used by the tests)
  b200msm_ctx* ctx = nullptr;
  int rc = b200msm_create(&ctx, device_ids[0]);
  if (rc) return rc;
  if (n_devices > 1) {
    ctx->devs.push_back(ctx);
    for (int i = 1; i < n_devices; i++) {
      b200msm_ctx* c = nullptr; rc = b200msm_create(&c, device_ids[i]);
      if (rc) { b200msm_destroy(ctx); return rc; }
      ctx->devs.push_back(c);
    }
    cudaSetDevice(ctx->device);
  }
  *out = ctx;
  return B200MSM_OK;
}
int b200msm_device_count(const b200msm_ctx* ctx) { return ctx ? (int)std::max<size_t>(1, ctx->devs.size()) : 0; }

// =================================================================== the MSM entry points
int b200msm_g1_multiexp_affine(b200msm_ctx* ctx, int curve, const void* bases, const void* scalars, uint32_t scalar_size, uint64_t n, void* out) {
  if (ctx && ctx->devs.size() > 1) return multi_msm(ctx, curve, bases, false, scalars, scalar_size, n, 0, 8 * scalar_size, out);
  return msm_entry(ctx, curve, bases, false, scalars, scalar_size, n, 0, 8 * scalar_size, out, nullptr);
}

int b200msm_g1_multiexp_affine_chunk(b200msm_ctx* ctx, int curve, const void* bases, const void* scalars, uint32_t scalar_size, uint64_t n,
                                     uint32_t start_bit, uint32_t chunk_bits, void* out) {
  if (ctx && (chunk_bits == 0 || chunk_bits > 32)) { ctx->err = "chunk_bits must be in [1, 32]"; return B200MSM_E_ARG; }
  if (ctx && ctx->devs.size() > 1) return multi_msm(ctx, curve, bases, false, scalars, scalar_size, n, start_bit, chunk_bits, out);
  return msm_entry(ctx, curve, bases, false, scalars, scalar_size, n, start_bit, chunk_bits, out, nullptr);
}

int b200msm_g1_multiexp(b200msm_ctx* ctx, int curve, const void* bases_jac, const void* scalars, uint32_t scalar_size, uint64_t n, void* out) {
  if (ctx && ctx->devs.size() > 1) return multi_msm(ctx, curve, bases_jac, true, scalars, scalar_size, n, 0, 8 * scalar_size, out);
  return msm_jacobian(ctx, curve, bases_jac, scalars, scalar_size, n, 0, 8 * scalar_size, out);
}
int b200msm_g1_multiexp_chunk(b200msm_ctx* ctx, int curve, const void* bases_jac, const void* scalars, uint32_t scalar_size, uint64_t n,
                              uint32_t start_bit, uint32_t chunk_bits, void* out) {
  if (ctx && (chunk_bits == 0 || chunk_bits > 32)) { ctx->err = "chunk_bits must be in [1, 32]"; return B200MSM_E_ARG; }
  if (ctx && ctx->devs.size() > 1) return multi_msm(ctx, curve, bases_jac, true, scalars, scalar_size, n, start_bit, chunk_bits, out);
  return msm_jacobian(ctx, curve, bases_jac, scalars, scalar_size, n, start_bit, chunk_bits, out);
}

int b200msm_upload_bases(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, uint64_t* handle) {
  if (!ctx) return B200MSM_E_ARG;
  if (ctx->devs.size() > 1) return multi_upload(ctx, curve, bases, n, false, 0, 0, handle);
  return single_upload(ctx, curve, bases, n, handle);
}
int b200msm_upload_bases_windowed(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, uint32_t scalar_size, uint32_t window_bits, uint64_t* handle) {
  if (!ctx) return B200MSM_E_ARG;
  if (ctx->devs.size() > 1) { if (!n || !bases || scalar_size == 0 || scalar_size > 32 || window_bits > 24) return B200MSM_E_ARG; return multi_upload(ctx, curve, bases, n, true, scalar_size, window_bits, handle); }
  return single_upload_windowed(ctx, curve, bases, n, scalar_size, window_bits, handle);
}
int b200msm_free_bases(b200msm_ctx* ctx, uint64_t handle) {
  if (!ctx) return B200MSM_E_ARG;
  auto it = ctx->mres.find(handle);
  if (it != ctx->mres.end()) {
    for (size_t g = 0; g < ctx->devs.size(); g++) if (it->second.h[g]) single_free(ctx->devs[g], it->second.h[g]);
    ctx->mres.erase(it); cudaSetDevice(ctx->device);
    return B200MSM_OK;
  }
  return single_free(ctx, handle);
}
int b200msm_g1_multiexp_resident(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size, uint64_t n, void* out, b200msm_stats* stats) {
  if (!ctx) return B200MSM_E_ARG;
  if (ctx->mres.count(handle)) return multi_resident(ctx, handle, scalars, scalar_size, n, out, stats);
  return single_resident(ctx, handle, 0, scalars, scalar_size, n, out, stats);
}

// count independent MSMs over the same resident bases (see single_batch / multi_batch)
int b200msm_g1_multiexp_batch(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size, uint64_t n, uint32_t count, void* out) {
  if (!ctx) return B200MSM_E_ARG;
  if (count == 0) return B200MSM_OK;
  if (ctx->mres.count(handle)) return multi_batch(ctx, handle, scalars, scalar_size, n, count, out);
  return single_batch(ctx, handle, scalars, scalar_size, n, count, 0, 1, out);
}

int b200msm_g1_normalize(b200msm_ctx* ctx, int curve, const void* jac, uint64_t count, void* xy) {
  if (!ctx || !curve_ok(curve) || !jac || !xy) return B200MSM_E_ARG;
  if (count == 0) return B200MSM_OK;
  CK(cudaSetDevice(ctx->device));
  const int n8 = n8_of(curve);
  const void* d_in; int rc = stage(ctx, jac, (size_t)count * 3 * n8, ctx->misc, &d_in); if (rc) return rc;
  CK(ctx->acc_e.ensure((size_t)count * 2 * n8));
  uint32_t g = (uint32_t)((count + 63) / 64);
  B200_CURVE_SWITCH(curve, k_normalize<C><<<g, 64, 0, ctx->stream>>>(d_in, ctx->acc_e.p, (uint32_t)count))
  CKL();
  return deliver(ctx, ctx->acc_e.p, xy, (size_t)count * 2 * n8);
}

int b200msm_g1_sum(b200msm_ctx* ctx, int curve, const void* pts, uint64_t count, void* out) {
  if (!ctx || !curve_ok(curve) || !out || (count && !pts)) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const int n8 = n8_of(curve);
  const void* d_in; int rc = stage(ctx, pts, (size_t)count * 3 * n8, ctx->misc, &d_in); if (rc) return rc;
  CK(ctx->out.ensure(3 * 96));
  B200_CURVE_SWITCH(curve, k_sum_jacobian<C><<<1, 32, 0, ctx->stream>>>(d_in, (uint32_t)count, ctx->out.p))
  CKL();
  return deliver(ctx, ctx->out.p, out, 3 * n8);
}

int b200msm_g1_generate_bases(b200msm_ctx* ctx, int curve, uint64_t seed, uint64_t first, uint64_t n, void* device_out) {
  if (!ctx || !curve_ok(curve) || (n && !is_device_ptr(device_out)) || n >= (1ull << 31)) return B200MSM_E_ARG;
  if (n == 0) return B200MSM_OK;
  CK(cudaSetDevice(ctx->device));
  const int n8 = n8_of(curve);
  // generator in affine Montgomery form (build_bls12381.js:99-111, build_bn128.js:58-70)
  static const uint32_t G_BLS[24] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u, 0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u,
                                     0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u, 0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
  static const uint32_t G_BN[16] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u,
                                    0x8b1e1b3au, 0xa6ba871bu, 0xeb8e167bu, 0x14f1d651u, 0xf0f28c58u, 0xccdd46deu, 0x340fbe5eu, 0x1c14ef83u};
  // G2 generators, x = x0 + x1 u, y = y0 + y1 u as x0 || x1 || y0 || y1 (build_bls12381.js:127-148, build_bn128.js:122-144)
  static const uint32_t G2_BLS[48] = {
    0x02940a10u, 0xf5f28fa2u, 0x87b4961au, 0xb3f5fb26u, 0x3e2ae580u, 0xa1a893b5u, 0x1a3caee9u, 0x9894999du, 0x1863366bu, 0x6f67b763u, 0x4350bcd7u, 0x05819192u,
    0x9e23f606u, 0xa5a9c075u, 0xbccd60c3u, 0xaaa0c59du, 0xe2867806u, 0x3bb17e18u, 0x8541b367u, 0x1b1ab6ccu, 0xf2158547u, 0xc2b6ed0eu, 0x7360edf3u, 0x11922a09u,
    0x60494c4au, 0x4c730af8u, 0x5e369c5au, 0x597cfa1fu, 0xaa0a635au, 0xe7e6856cu, 0x6e0d495fu, 0xbbefb5e9u, 0xf0ef25a2u, 0x07d3a975u, 0x7e80dae5u, 0x0083fd8eu,
    0xdf64b05du, 0xadc0fc92u, 0x2b1461dcu, 0x18aa270au, 0x3be4eba0u, 0x86adac6au, 0xc93da33au, 0x79495c4eu, 0xa43ccaedu, 0xe7175850u, 0x63de1bf2u, 0x0b2bc2a1u};
  static const uint32_t G2_BN[32] = {
    0x02bc2026u, 0x8e83b5d1u, 0x497b0172u, 0xdceb1935u, 0x97811adfu, 0xfbb82647u, 0xaf96503bu, 0x19573841u,
    0xa84c6140u, 0xafb4737du, 0x5802d8c4u, 0x6043dd5au, 0x52a02f86u, 0x09e950fcu, 0x3aea7b6bu, 0x14fef083u,
    0x886be9f6u, 0x619dfa9du, 0xf59e9b78u, 0xfe7fd297u, 0x231b7dfeu, 0xff9e1a62u, 0xae9e4206u, 0x28fd7eebu,
    0xc71856eeu, 0x64095b56u, 0x327d3cbbu, 0xdc57f922u, 0x33351076u, 0x55f935beu, 0x93fd6482u, 0x0da4a0e6u};
  const void* gens[4] = {G_BLS, G_BN, G2_BLS, G2_BN};
  CK(ctx->misc.ensure(256));
  CK(cudaMemcpyAsync(ctx->misc.p, gens[curve], 2 * n8, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx->acc_a.ensure((size_t)n * 4 * n8));
  uint32_t g = (uint32_t)((n + 127) / 128);
  constexpr int GROUP = 16;
  uint32_t g2 = (uint32_t)(((n + GROUP - 1) / GROUP + 127) / 128);
  B200_CURVE_SWITCH(curve,
    k_generate_xyzz<C><<<g, 128, 0, ctx->stream>>>(ctx->misc.p, seed, first, (uint32_t)n, ctx->acc_a.p); CKL();
    k_xyzz_to_affine<C, GROUP><<<g2, 128, 0, ctx->stream>>>(ctx->acc_a.p, (uint32_t)n, device_out); CKL())
  CK(cudaStreamSynchronize(ctx->stream));
  return B200MSM_OK;
}

int b200msm_g1_batch_convert(b200msm_ctx* ctx, int curve, int op, const void* in, uint64_t n, void* out) {
  if (!ctx || !curve_ok(curve) || op < 0 || op > 5 || (n && (!in || !out)) || n >= (1ull << 31)) return B200MSM_E_ARG;
  if (n == 0) return B200MSM_OK;
  CK(cudaSetDevice(ctx->device));
  const size_t n8 = n8_of(curve);
  const size_t in_sz[6] = {2 * n8, 2 * n8, 2 * n8, n8, 3 * n8, 2 * n8}, out_sz[6] = {2 * n8, n8, 2 * n8, 2 * n8, 2 * n8, 3 * n8};
  const void* d_in; int rc = stage(ctx, in, n * in_sz[op], ctx->acc_a, &d_in); if (rc) return rc;
  CK(ctx->acc_c.ensure(n * out_sz[op] + 16));
  const uint32_t g = (uint32_t)((n + 127) / 128);
  if (op == CODEC_TO_AFFINE) {
    constexpr int GROUP = 8;
    const uint32_t g2 = (uint32_t)(((n + GROUP - 1) / GROUP + 127) / 128);
    B200_CURVE_SWITCH(curve, k_jacobian_to_affine<C, GROUP><<<g2, 128, 0, ctx->stream>>>((const uint8_t*)d_in, (uint32_t)n, ctx->acc_c.as<uint8_t>()))
  } else { B200_CURVE_SWITCH(curve, k_codec<C><<<g, 128, 0, ctx->stream>>>(op, (const uint8_t*)d_in, (uint32_t)n, ctx->acc_c.as<uint8_t>())) }
  CKL();
  return deliver(ctx, ctx->acc_c.p, out, n * out_sz[op]);
}

// g1m_glv_decomposeScalar over a batch (build_glv.js:53-146): out_scalars n x 64 bytes, out_signs n x u32 (nullable)
int b200msm_glv_decompose_scalars(b200msm_ctx* ctx, int curve, const void* scalars, uint64_t n, void* out_scalars, void* out_signs) {
  if (!ctx || (n && (!scalars || !out_scalars)) || n >= (1ull << 31)) return B200MSM_E_ARG;
  if (curve != B200MSM_BLS12_381_G1) { ctx->err = "GLV constants exist for BLS12-381 only (build_glv.js:3)"; return B200MSM_E_UNSUPPORTED; }
  if (n == 0) return B200MSM_OK;
  CK(cudaSetDevice(ctx->device));
  const void* d_s; int rc = stage(ctx, scalars, n * 32, ctx->acc_a, &d_s); if (rc) return rc;
  CK(ctx->acc_c.ensure(n * 64 + 16)); CK(ctx->acc_d.ensure(n * 4 + 16));
  k_glv_decompose<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t*)d_s, (uint32_t)n, ctx->acc_c.as<uint32_t>(), ctx->acc_d.as<uint32_t>()); CKL();
  rc = deliver(ctx, ctx->acc_c.p, out_scalars, n * 64); if (rc) return rc;
  if (out_signs) return deliver(ctx, ctx->acc_d.p, out_signs, n * 4);
  return B200MSM_OK;
}

// g1m_glv_preprocessEndomorphism (build_glv.js:178-263): n points / n scalars -> 2n points / 2n scalars of 32 bytes (< 2^128)
int b200msm_g1_glv_preprocess(b200msm_ctx* ctx, int curve, const void* points, const void* scalars, uint64_t n, void* out_points, void* out_scalars) {
  if (!ctx || (n && (!points || !scalars || !out_points || !out_scalars)) || n >= (1ull << 30)) return B200MSM_E_ARG;
  if (curve != B200MSM_BLS12_381_G1) { ctx->err = "GLV constants exist for BLS12-381 only (build_glv.js:3)"; return B200MSM_E_UNSUPPORTED; }
  if (n == 0) return B200MSM_OK;
  CK(cudaSetDevice(ctx->device));
  const void *d_s, *d_p; int rc = stage(ctx, scalars, n * 32, ctx->acc_a, &d_s); if (rc) return rc;
  rc = stage(ctx, points, n * 96, ctx->acc_b, &d_p); if (rc) return rc;
  const bool so_dev = is_device_ptr(out_scalars) && (reinterpret_cast<uintptr_t>(out_scalars) & 15) == 0;
  const bool po_dev = is_device_ptr(out_points) && (reinterpret_cast<uintptr_t>(out_points) & 15) == 0;
  if (!so_dev) CK(ctx->acc_c.ensure(n * 64 + 16));
  if (!po_dev) CK(ctx->acc_e.ensure(n * 192 + 16));
  CK(ctx->acc_d.ensure(n * 4 + 16));
  uint32_t* d_so = so_dev ? (uint32_t*)out_scalars : ctx->acc_c.as<uint32_t>();
  void* d_po = po_dev ? out_points : ctx->acc_e.p;
  const uint32_t g = (uint32_t)((n + 127) / 128);
  k_glv_decompose<<<g, 128, 0, ctx->stream>>>((const uint32_t*)d_s, (uint32_t)n, d_so, ctx->acc_d.as<uint32_t>()); CKL();
  k_glv_points<<<g, 128, 0, ctx->stream>>>(d_p, ctx->acc_d.as<uint32_t>(), (uint32_t)n, d_po); CKL();
  if (!so_dev) { rc = deliver(ctx, d_so, out_scalars, n * 64); if (rc) return rc; }
  if (!po_dev) { rc = deliver(ctx, d_po, out_points, n * 192); if (rc) return rc; }
  return B200MSM_OK;
}

// f1m_batchInverse (build_batchinverse.js:4-140): out[i] = 1 / in[i], zeros stay zero (the reference skips them the same way), through the
// product tree the batch-affine rounds use.  Elements are Montgomery residues of the curve's coordinate field (Fq, or Fq2 for the G2 ids).
extern "C++" {
namespace {
template <class C> __global__ void k_binv_prepare(const void* __restrict__ in, void* __restrict__ vals, uint8_t* __restrict__ zero, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  Fe<C::N> v; fe_load_cg<C>(v, reinterpret_cast<const char*>(in) + (uint64_t)i * 4 * C::N);
  const bool z = fe_is_zero<C>(v); zero[i] = z;
  if (z) fe_set_one<C>(v);
  fe_store<C>(reinterpret_cast<char*>(vals) + (uint64_t)i * 4 * C::N, v);
}
template <class C> __global__ void k_binv_finish(void* __restrict__ vals, const uint8_t* __restrict__ zero, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n || !zero[i]) return;
  Fe<C::N> v; fe_set_zero<C>(v); fe_store<C>(reinterpret_cast<char*>(vals) + (uint64_t)i * 4 * C::N, v);
}
}
}
int b200msm_fq_batch_inverse(b200msm_ctx* ctx, int curve, const void* in, uint64_t count, void* out) {
  if (!ctx || !curve_ok(curve) || !in || !out) return B200MSM_E_ARG;
  if (count == 0) return B200MSM_OK;
  if (count >= (1ull << 31)) { ctx->err = "count must be < 2^31"; return B200MSM_E_UNSUPPORTED; }
  CK(cudaSetDevice(ctx->device));
  const size_t n8 = (size_t)n8_of(curve), bytes = (size_t)count * n8;
  const void* din; int rc = stage(ctx, in, bytes, ctx->acc_a, &din); if (rc) return rc;
  TreeLane& ln = ctx->lane[0];
  CK(ln.prod.ensure((2 * count + 4096) * n8)); CK(ln.lvlprefix.ensure((2 * count + 4096) * n8)); CK(ln.others.ensure((WARP_LEVEL_MAX / 4 + 4096) * n8));
  CK(ctx->misc.ensure(std::max<size_t>(count, 2048)));
  const uint32_t g = (uint32_t)((count + 255) / 256);
  B200_CURVE_SWITCH(curve,
    k_binv_prepare<C><<<g, 256, 0, ctx->stream>>>(din, ln.prod.p, ctx->misc.as<uint8_t>(), (uint32_t)count); CKL();
    rc = product_tree_invert<C>(ctx, ln, ctx->stream, ln.prod.as<char>(), ln.lvlprefix.as<char>(), count, ctx->opt_pt_k);
    if (!rc) { k_binv_finish<C><<<g, 256, 0, ctx->stream>>>(ln.prod.p, ctx->misc.as<uint8_t>(), (uint32_t)count); CKL(); })
  if (rc) return rc;
  return deliver(ctx, ln.prod.p, out, bytes);
}

// Test hook: the digit / sort phase alone -- signed-digit recoding + histogram (k_digits<count>), exclusive scan, scatter (k_digits<scatter>)
// (computeSchedule + organizeBuckets, build_multiexp_opt.js:175-633).  plan_out = {Wd windows, W bucket slots, B buckets per slot, c0, rem, nbits}:
// windows 0 .. Wd-1 are c0 + 1 bits wide for the first `rem` of them and c0 bits after that; offsets_out (host, W*B + 1 words) = segment starts of
// every bucket (slot-major), sorted_out (host, >= offsets[W*B] words) = point index | sign << 31 of every pair, bucket by bucket.
int b200msm_debug_schedule(b200msm_ctx* ctx, const void* scalars, uint32_t scalar_size, uint64_t n64, uint32_t window_bits, uint32_t plan_out[6],
                           uint32_t* offsets_out, uint64_t offsets_cap, uint32_t* sorted_out, uint64_t sorted_cap) {
  if (!ctx || !scalars || !plan_out || scalar_size == 0 || scalar_size > 32 || n64 == 0 || n64 >= (1ull << 31)) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream; const uint32_t n = (uint32_t)n64, nbits = 8 * scalar_size;
  const void* d_sraw; int rc = stage(ctx, scalars, (size_t)n * scalar_size, ctx->scalars, &d_sraw); if (rc) return rc;
  CK(ctx->canon.ensure((size_t)n * 32));
  k_canon_scalars<<<(n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_sraw), scalar_size, n, 0, nbits, ctx->canon.as<uint32_t>()); CKL();
  MsmPlan pl; pl.n = n; pl.nbits = nbits; pl.pre_stride = 0;
  const uint32_t ct = window_bits > 0 ? std::min<uint32_t>(window_bits, std::min<uint32_t>(nbits, 24)) : auto_window_bits(n, nbits);
  pl.Wd = (nbits + ct - 1) / ct; pl.c0 = nbits / pl.Wd; pl.rem = nbits - pl.c0 * pl.Wd;
  pl.c = pl.c0 + (pl.rem ? 1 : 0); pl.B = 1u << (pl.c - 1); pl.logB = pl.c - 1; pl.W = pl.Wd + (pl.rem == 0 ? 1 : 0);
  const uint64_t nb = (uint64_t)pl.W * pl.B;
  if (nb > (1ull << 28) || (uint64_t)n * pl.W >= (1ull << 32)) { ctx->err = "schedule too large for the test hook"; return B200MSM_E_UNSUPPORTED; }
  plan_out[0] = pl.Wd; plan_out[1] = pl.W; plan_out[2] = pl.B; plan_out[3] = pl.c0; plan_out[4] = pl.rem; plan_out[5] = nbits;
  if (!offsets_out || offsets_cap < nb + 1) return offsets_out ? B200MSM_E_ARG : B200MSM_OK;       // plan only
  CK(ctx->counts.ensure(nb * 4)); CK(ctx->offsets.ensure((nb + 1 + 512) * 4)); CK(ctx->sorted.ensure((size_t)n * pl.W * 4 + 16)); CK(ctx->ranks.ensure((size_t)n * pl.Wd * 4 + 16));
  CK(cudaMemsetAsync(ctx->counts.p, 0, nb * 4, s));
  const uint32_t tb = 256, gb = (n + tb - 1) / tb;
  k_digits<false><<<gb, tb, 0, s>>>(ctx->canon.as<uint32_t>(), pl, ctx->counts.as<uint32_t>(), nullptr, ctx->ranks.as<uint32_t>(), nullptr, 0u, pl.Wd); CKL();
  rc = exclusive_scan(ctx, s, ctx->tiles, ctx->counts.as<uint32_t>(), ctx->offsets.as<uint32_t>(), (uint32_t)nb, 0u); if (rc) return rc;
  k_digits<true><<<gb, tb, 0, s>>>(ctx->canon.as<uint32_t>(), pl, nullptr, ctx->offsets.as<uint32_t>(), ctx->ranks.as<uint32_t>(), ctx->sorted.as<uint32_t>(), 0u, pl.Wd); CKL();
  CK(cudaMemcpyAsync(offsets_out, ctx->offsets.p, (nb + 1) * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const uint64_t pairs = offsets_out[nb];
  if (sorted_out) { if (sorted_cap < pairs) { ctx->err = "sorted_out too small"; return B200MSM_E_ARG; }
    CK(cudaMemcpyAsync(sorted_out, ctx->sorted.p, pairs * 4, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
  return B200MSM_OK;
}

int b200msm_fq_op(b200msm_ctx* ctx, int curve, int op, const void* a, const void* b, void* r, uint64_t count) {
  if (!ctx || !curve_ok(curve) || op < 0 || op > 10 || !a || !r) return B200MSM_E_ARG;
  if (count == 0) return B200MSM_OK;
  CK(cudaSetDevice(ctx->device));
  const int n8 = n8_of(curve); size_t bytes = (size_t)count * n8;
  const void *da, *db = nullptr; int rc;
  rc = stage(ctx, a, bytes, ctx->acc_a, &da); if (rc) return rc;
  if (b) { rc = stage(ctx, b, bytes, ctx->acc_b, &db); if (rc) return rc; }
  CK(ctx->acc_c.ensure(bytes));
  uint32_t g = (uint32_t)((count + 127) / 128);
  if (op >= 9) {
#if defined(B200_EXPERIMENTS)
    if (!db) { ctx->err = "op 9/10 need two operands"; return B200MSM_E_ARG; }
    if (!curve_g1(curve)) { ctx->err = "radix-2^29 multiplier: prime fields only"; return B200MSM_E_UNSUPPORTED; }
    if (curve == 0) k_fp29_mul<BLS12_381><<<g, 128, 0, ctx->stream>>>(da, db, ctx->acc_c.p, (uint32_t)count, op - 9);
    else k_fp29_mul<BN254><<<g, 128, 0, ctx->stream>>>(da, db, ctx->acc_c.p, (uint32_t)count, op - 9);
#else
    ctx->err = "ops 9/10 (radix-2^29 multiplier) exist only in -DB200_EXPERIMENTS builds"; return B200MSM_E_UNSUPPORTED;
#endif
  } else { B200_CURVE_SWITCH(curve, k_fp_op<C><<<g, 128, 0, ctx->stream>>>(op, da, db, ctx->acc_c.p, (uint32_t)count)) }
  CKL();
  return deliver(ctx, ctx->acc_c.p, r, bytes);
}

// kind 0: IMAD.WIDE.U32 (32x32+64 with carry, the multiplier's instruction); kind 1: 32-bit IMAD
static int probe_int_pipe(b200msm_ctx* ctx, int kind, double* per_s) {
  if (!ctx || !per_s) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
  CK(ctx->misc.ensure(256));
  const uint32_t blocks = prop.multiProcessorCount * 8, iters = 4096;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (kind == 0) k_imadx_probe<<<blocks, 256, 0, ctx->stream>>>(iters, 12345u + rep, ctx->misc.as<uint32_t>());
    else k_imad_probe<1><<<blocks, 256, 0, ctx->stream>>>(iters, 12345u + rep, ctx->misc.as<unsigned long long>());
    CKL();
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); CK(cudaEventSynchronize(ctx->ev[1]));
    float ms; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    double rate = (double)blocks * 256.0 * iters * (kind == 0 ? 16.0 : 64.0) / (ms * 1e-3);
    if (rep && rate > best) best = rate;
  }
  *per_s = best; return B200MSM_OK;
}
int b200msm_probe_imad(b200msm_ctx* ctx, double* imad_wide_per_s) { return probe_int_pipe(ctx, 0, imad_wide_per_s); }
int b200msm_probe_imad32(b200msm_ctx* ctx, double* imad32_per_s) { return probe_int_pipe(ctx, 1, imad32_per_s); }

int b200msm_probe_fqmul(b200msm_ctx* ctx, int curve, double* fqmul_per_s) {
  if (!ctx || !fqmul_per_s || !curve_ok(curve)) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
  const uint32_t blocks = prop.multiProcessorCount * 4, threads = 256, iters = 512 | (ctx->probe_sqr ? 0x80000000u : 0u);
  const int n8 = n8_of(curve);
  CK(ctx->acc_a.ensure(1024 * 96)); CK(ctx->acc_b.ensure((size_t)blocks * threads * n8));
  CK(cudaMemsetAsync(ctx->acc_a.p, 0x17, 1024 * 96, ctx->stream));
  if (!curve_g1(curve) && (ctx->probe29 || ctx->opt_probe_smem)) { ctx->err = "probe29 / probe_smem: prime fields only"; return B200MSM_E_UNSUPPORTED; }
  const size_t psm = (size_t)ctx->opt_probe_smem;       // dynamic shared memory per block, only to cap the resident warps per SM (occupancy sensitivity of the multiplier)
  if (psm) { CK(cudaFuncSetAttribute(k_fpmul_probe<BLS12_381>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); CK(cudaFuncSetAttribute(k_fpmul_probe<BN254>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); }
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
#if defined(B200_EXPERIMENTS)
    if (ctx->probe29) {
      if (curve == 0) k_fpmul29_probe<BLS12_381><<<blocks, threads, 0, ctx->stream>>>(iters, ctx->acc_a.p, ctx->acc_b.p);
      else k_fpmul29_probe<BN254><<<blocks, threads, 0, ctx->stream>>>(iters, ctx->acc_a.p, ctx->acc_b.p);
    } else
#endif
    { B200_CURVE_SWITCH(curve, k_fpmul_probe<C><<<blocks, threads, psm, ctx->stream>>>(iters, ctx->acc_a.p, ctx->acc_b.p)) }
    CKL();
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); CK(cudaEventSynchronize(ctx->ev[1]));
    float ms; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    double rate = (double)blocks * threads * (iters & 0x7fffffffu) / (ms * 1e-3);
    if (rep && rate > best) best = rate;
  }
  *fqmul_per_s = best; return B200MSM_OK;
}

#if defined(B200_EXPERIMENTS)
int b200msm_probe_dfma(b200msm_ctx* ctx, double* dfma_per_s) {
  if (!ctx || !dfma_per_s) return B200MSM_E_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
  const uint32_t blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  CK(ctx->acc_b.ensure((size_t)blocks * threads * 8));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    k_dfma_probe<<<blocks, threads, 0, ctx->stream>>>(iters, 1.000001, ctx->acc_b.as<double>()); CKL();
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); CK(cudaEventSynchronize(ctx->ev[1]));
    float ms; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    double rate = (double)blocks * threads * iters * 8 / (ms * 1e-3);
    if (rep && rate > best) best = rate;
  }
  *dfma_per_s = best; return B200MSM_OK;
}

#endif  // B200_EXPERIMENTS

int b200msm_get_counter(b200msm_ctx* ctx, const char* key, uint64_t* value) {
  if (!ctx || !key || !value) return B200MSM_E_ARG;
  if (!strcmp(key, "launches")) { uint64_t t = ctx->launches; for (size_t g = 1; g < ctx->devs.size(); g++) t += ctx->devs[g]->launches;      // all devices of a multi context
    for (b200msm_ctx* w : ctx->workers) t += w->launches; *value = t; return B200MSM_OK; }
  return B200MSM_E_ARG;
}

int b200msm_constants(int curve, uint32_t* n8, uint8_t* q, uint8_t* r_mod_q, uint8_t* r2_mod_q, uint32_t* np32) {
  if (!curve_ok(curve)) return B200MSM_E_ARG;
  auto put = [](uint8_t* d, int i, uint32_t v) { if (d) memcpy(d + 4 * i, &v, 4); };
  // G2 ids describe their BASE field Fq (an Fq2 element is two of these): ids 0 and 2 -> BLS12-381, 1 and 3 -> BN254
  if ((curve & 1) == 0) {
    if (n8) *n8 = 48; if (np32) *np32 = BLS12_381::NP;
    for (int i = 0; i < BLS12_381::N; i++) { put(q, i, BLS12_381::q(i)); put(r_mod_q, i, BLS12_381::one(i)); put(r2_mod_q, i, BLS12_381::r2(i)); }
  } else {
    if (n8) *n8 = 32; if (np32) *np32 = BN254::NP;
    for (int i = 0; i < BN254::N; i++) { put(q, i, BN254::q(i)); put(r_mod_q, i, BN254::one(i)); put(r2_mod_q, i, BN254::r2(i)); }
  }
  return B200MSM_OK;
}

}  // extern "C"
