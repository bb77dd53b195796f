// ntt.cu -- number-theoretic transform over Fr on the GPU: the step on the other side of the MSM in a Groth16 / PLONK prover
// (SURVEY.md 8f row 4).  GPU form of frm_fft / frm_ifft, wasmcurves/src/build_fft.js:178-394 (+ __fftFinal :396-516,
// __reversePermutation :518-583): the reference bit-reverses in place, runs log2(n) radix-2 decimation-in-time stages with the
// twiddle of stage s taken from ROOTs[s], and for the inverse reverses the order (i -> n - i) and multiplies by 1/n.
// The result is the DFT  out[k] = sum_j in[j] * w^(j*k),  w = ROOTs[log2 n]  (resp. its inverse), so any exact algorithm gives
// the same canonical Montgomery bytes; parity is checked against the reference module's own frm_fft / frm_ifft exports.
//
// Here: the first 10 stages of every 1024-element tile in shared memory, reading the inputs from their bit-reversed positions (one launch),
// the remaining stages as radix-4 (two stages per pass over the data) or radix-2 global passes, twiddles w^e from a table of
// n/2 entries built once per (context, curve, size).  Fr elements are 32 bytes = one DRAM sector.
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <mutex>
#include <string>
#include "internal.h"
#include "fr.cuh"

using namespace b200;

namespace {

constexpr int TILE_LOG = 10, TILE = 1 << TILE_LOG;

// roots[s] = primitive 2^s-th root of unity (ROOTs, build_fft.js:44-63), s = 0..MAXBITS; aux[0] = 1/n (INV2[bits], :65-77)
template <class F>
__global__ void k_ntt_setup(void* __restrict__ roots, void* __restrict__ aux, uint32_t log2n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fe<F::N> w;
#pragma unroll
  for (int i = 0; i < F::N; i++) w.l[i] = F::root(i);
  for (int s = F::MAXBITS; s >= 0; s--) { fe_store<F>(reinterpret_cast<char*>(roots) + (size_t)s * 4 * F::N, w); fe_sqr<F>(w, w); }
  Fe<F::N> n, ninv;
#pragma unroll
  for (int i = 0; i < F::N; i++) n.l[i] = 0;
  n.l[log2n >> 5] = 1u << (log2n & 31);
  fe_to_mont<F>(n, n); fe_inv_fast<F>(ninv, n);
  fe_store<F>(aux, ninv);
}
// W[e] = w_n^e, e < n/2: product of the roots selected by the bits of e (w_n^(2^t) = roots[log2n - t])
template <class F>
__global__ void __launch_bounds__(256) k_ntt_twiddles(void* __restrict__ W, const void* __restrict__ roots, uint32_t log2n, int inverse) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (log2n == 0 || e >= (1u << (log2n - 1))) return;
  Fe<F::N> acc; fe_set_one<F>(acc);
  for (uint32_t t = 0; t + 1 < log2n; t++) if ((e >> t) & 1) {
    Fe<F::N> p; fe_load<F>(p, reinterpret_cast<const char*>(roots) + (size_t)(log2n - t) * 4 * F::N);
    fe_mul<F>(acc, acc, p);
  }
  (void)inverse;
  fe_store<F>(reinterpret_cast<char*>(W) + (size_t)e * 4 * F::N, acc);
}
// __reversePermutation as a gather: out[i] = in[bitrev(i)]
template <class F>
__global__ void __launch_bounds__(256) k_ntt_bitrev(const void* __restrict__ in, void* __restrict__ out, uint32_t log2n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << log2n)) return;
  const uint32_t j = log2n ? (__brev(i) >> (32 - log2n)) : 0;
  Fe<F::N> v; fe_load<F>(v, reinterpret_cast<const char*>(in) + (size_t)j * 4 * F::N);
  fe_store<F>(reinterpret_cast<char*>(out) + (size_t)i * 4 * F::N, v);
}
// stages 1 .. K (K = min(log2n, 10)) of every tile of 2^K consecutive elements, in shared memory (limb-major: conflict-free)
// two consecutive radix-2 stages s and s + 1 on the four elements i0, i0 + h, i0 + 2h, i0 + 3h of one block of 4h (h = 2^(s-1), j = i0 mod h)
template <class F>
__device__ __forceinline__ void butterfly4(Fe<F::N>& a0, Fe<F::N>& a1, Fe<F::N>& a2, Fe<F::N>& a3, const char* __restrict__ tw, uint32_t log2n, uint32_t s, uint32_t j, uint32_t h) {
  Fe<F::N> w, t;
  // stage s: pairs (a0, a1) and (a2, a3), both with twiddle w_(2^s)^j
  fe_load<F>(w, tw + ((size_t)j << (log2n - s)) * 4 * F::N);
  fe_mul<F>(t, a1, w); fe_sub<F>(a1, a0, t); fe_add<F>(a0, a0, t);
  fe_mul<F>(t, a3, w); fe_sub<F>(a3, a2, t); fe_add<F>(a2, a2, t);
  // stage s + 1: pairs (a0, a2) with w_(2^(s+1))^j and (a1, a3) with w_(2^(s+1))^(j + h)
  fe_load<F>(w, tw + ((size_t)j << (log2n - s - 1)) * 4 * F::N);
  fe_mul<F>(t, a2, w); fe_sub<F>(a2, a0, t); fe_add<F>(a0, a0, t);
  fe_load<F>(w, tw + ((size_t)(j + h) << (log2n - s - 1)) * 4 * F::N);
  fe_mul<F>(t, a3, w); fe_sub<F>(a3, a1, t); fe_add<F>(a1, a1, t);
}

// The tile reads its inputs straight from their bit-reversed positions (__reversePermutation fused into the load: element i of tile b
// is in[bitrev(b*T + i)], a gather of whole 32-byte sectors), so the permuted array is never written out and read back.
template <class F>
__global__ void __launch_bounds__(TILE / 4) k_ntt_tile(const void* __restrict__ in, void* __restrict__ x, const void* __restrict__ W, uint32_t log2n, uint32_t K) {
  __shared__ uint32_t sm[F::N][TILE];
  const uint32_t T = 1u << K, half_threads = T >> 1;
  char* base = reinterpret_cast<char*>(x) + (size_t)blockIdx.x * T * 4 * F::N;
  for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
    const uint32_t src = __brev(blockIdx.x * T + i) >> (32 - log2n);
    Fe<F::N> v; fe_load<F>(v, reinterpret_cast<const char*>(in) + (size_t)src * 4 * F::N);
#pragma unroll
    for (int k = 0; k < F::N; k++) sm[k][i] = v.l[k];
  }
  __syncthreads();
  uint32_t s = 1;
  if (K & 1) {                                   // odd number of stages: stage 1 alone (all its twiddles are 1)
    for (uint32_t b = threadIdx.x; b < half_threads; b += blockDim.x) {
      const uint32_t i1 = 2 * b, i2 = i1 + 1;
      Fe<F::N> u, v, w;
#pragma unroll
      for (int k = 0; k < F::N; k++) { u.l[k] = sm[k][i1]; v.l[k] = sm[k][i2]; }
      fe_add<F>(w, u, v); fe_sub<F>(v, u, v);
#pragma unroll
      for (int k = 0; k < F::N; k++) { sm[k][i1] = w.l[k]; sm[k][i2] = v.l[k]; }
    }
    __syncthreads();
    s = 2;
  }
  for (; s + 1 <= K; s += 2) {                   // two stages per barrier: each thread owns the four elements of a radix-4 butterfly
    const uint32_t h = 1u << (s - 1);
    for (uint32_t b = threadIdx.x; b < (T >> 2); b += blockDim.x) {
      const uint32_t j = b & (h - 1), i0 = ((b >> (s - 1)) << (s + 1)) + j;
      Fe<F::N> a0, a1, a2, a3;
#pragma unroll
      for (int k = 0; k < F::N; k++) { a0.l[k] = sm[k][i0]; a1.l[k] = sm[k][i0 + h]; a2.l[k] = sm[k][i0 + 2 * h]; a3.l[k] = sm[k][i0 + 3 * h]; }
      butterfly4<F>(a0, a1, a2, a3, reinterpret_cast<const char*>(W), log2n, s, j, h);
#pragma unroll
      for (int k = 0; k < F::N; k++) { sm[k][i0] = a0.l[k]; sm[k][i0 + h] = a1.l[k]; sm[k][i0 + 2 * h] = a2.l[k]; sm[k][i0 + 3 * h] = a3.l[k]; }
    }
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
    Fe<F::N> v;
#pragma unroll
    for (int k = 0; k < F::N; k++) v.l[k] = sm[k][i];
    fe_store<F>(base + (size_t)i * 4 * F::N, v);
  }
}
// one radix-2 stage s over the whole array: one butterfly per thread
template <class F>
__global__ void __launch_bounds__(256) k_ntt_stage2(void* __restrict__ x, const void* __restrict__ W, uint32_t log2n, uint32_t s) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (1u << (log2n - 1))) return;
  const uint32_t half = 1u << (s - 1), j = b & (half - 1), i1 = ((b >> (s - 1)) << s) + j, i2 = i1 + half;
  char* p = reinterpret_cast<char*>(x);
  Fe<F::N> u, v, w, t;
  fe_load_cg<F>(u, p + (size_t)i1 * 4 * F::N); fe_load_cg<F>(v, p + (size_t)i2 * 4 * F::N);
  fe_load<F>(w, reinterpret_cast<const char*>(W) + ((size_t)j << (log2n - s)) * 4 * F::N);
  fe_mul<F>(t, v, w);
  fe_add<F>(v, u, t); fe_sub<F>(w, u, t);
  fe_store<F>(p + (size_t)i1 * 4 * F::N, v); fe_store<F>(p + (size_t)i2 * 4 * F::N, w);
}
// stages s and s+1 in one pass: each thread owns the four elements i, i + h, i + 2h, i + 3h (h = 2^(s-1)) of one block of 4h
template <class F>
__global__ void __launch_bounds__(256) k_ntt_stage4(void* __restrict__ x, const void* __restrict__ W, uint32_t log2n, uint32_t s) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (1u << (log2n - 2))) return;
  const uint32_t h = 1u << (s - 1), j = b & (h - 1), i0 = ((b >> (s - 1)) << (s + 1)) + j;
  char* p = reinterpret_cast<char*>(x);
  const char* tw = reinterpret_cast<const char*>(W);
  Fe<F::N> a0, a1, a2, a3;
  fe_load_cg<F>(a0, p + (size_t)i0 * 4 * F::N); fe_load_cg<F>(a1, p + (size_t)(i0 + h) * 4 * F::N);
  fe_load_cg<F>(a2, p + (size_t)(i0 + 2 * h) * 4 * F::N); fe_load_cg<F>(a3, p + (size_t)(i0 + 3 * h) * 4 * F::N);
  butterfly4<F>(a0, a1, a2, a3, tw, log2n, s, j, h);
  fe_store<F>(p + (size_t)i0 * 4 * F::N, a0); fe_store<F>(p + (size_t)(i0 + h) * 4 * F::N, a1);
  fe_store<F>(p + (size_t)(i0 + 2 * h) * 4 * F::N, a2); fe_store<F>(p + (size_t)(i0 + 3 * h) * 4 * F::N, a3);
}
// __fftFinal for the inverse (build_fft.js:396-516): x[i] <-> x[n - i], everything times 1/n
template <class F>
__global__ void __launch_bounds__(256) k_ntt_final_inverse(void* __restrict__ x, const void* __restrict__ aux, uint32_t log2n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = 1u << log2n;
  if (i > n / 2) return;
  char* p = reinterpret_cast<char*>(x);
  Fe<F::N> f, a, b; fe_load<F>(f, aux);
  const uint32_t k = (n - i) & (n - 1);
  fe_load_cg<F>(a, p + (size_t)i * 4 * F::N);
  if (k == i) { fe_mul<F>(a, a, f); fe_store<F>(p + (size_t)i * 4 * F::N, a); return; }
  fe_load_cg<F>(b, p + (size_t)k * 4 * F::N);
  fe_mul<F>(a, a, f); fe_mul<F>(b, b, f);
  fe_store<F>(p + (size_t)i * 4 * F::N, b); fe_store<F>(p + (size_t)k * 4 * F::N, a);
}

struct NttState {
  void* roots[2] = {nullptr, nullptr};         // per curve: (MAXBITS + 1) roots + 1/n slot
  uint32_t setup_log2n[2] = {~0u, ~0u};        // the size the roots' 1/n slot was computed for (the setup kernel is a serial chain of ~32 squarings + one inversion:
                                               // run once per (curve, size), not once per transform)
  void* W[2] = {nullptr, nullptr}; uint32_t W_log2n[2] = {0, 0}; size_t W_cap[2] = {0, 0};
  void* buf = nullptr; size_t buf_cap = 0;     // staging / ping buffer
  void* buf2 = nullptr; size_t buf2_cap = 0;
  cudaEvent_t ev[5] = {};                      // phase boundaries of the last transform: start | bit reversal | tile stages | radix-4/2 passes | final
  uint32_t last_passes4 = 0, last_passes2 = 0;
};
std::map<b200msm_ctx*, NttState> g_state;
std::mutex g_mu;

bool is_dev(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

#define NCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { b200msm_internal_set_error(ctx, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); \
  return e_ == cudaErrorMemoryAllocation ? B200MSM_E_NOMEM : B200MSM_E_CUDA; } } while (0)

int ensure(b200msm_ctx* ctx, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return B200MSM_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  NCK(cudaMalloc(p, bytes + 256)); *cap = bytes + 256; return B200MSM_OK;
}

template <class F>
int run_ntt(b200msm_ctx* ctx, NttState& st, int ci, const void* in, uint32_t L, int inverse, void* out) {
  cudaStream_t s = b200msm_internal_stream(ctx);
  const size_t fe = 4 * F::N, n = (size_t)1 << L, bytes = n * fe;
  uint64_t launches = 0;
  if (!st.roots[ci]) NCK(cudaMalloc(&st.roots[ci], (F::MAXBITS + 2) * fe));
  char* aux = reinterpret_cast<char*>(st.roots[ci]) + (size_t)(F::MAXBITS + 1) * fe;
  if (st.setup_log2n[ci] != L) { k_ntt_setup<F><<<1, 32, 0, s>>>(st.roots[ci], aux, L); launches++; st.setup_log2n[ci] = L; }
  // (tables built on the stream of an earlier call stay valid across b200msm_set_stream, which drains the old stream before switching)
  if (L >= 1 && st.W_log2n[ci] != L) {
    int rc = ensure(ctx, &st.W[ci], &st.W_cap[ci], (n / 2 + 1) * fe); if (rc) return rc;
    const uint32_t cnt = (uint32_t)(n / 2);
    k_ntt_twiddles<F><<<(cnt + 255) / 256, 256, 0, s>>>(st.W[ci], st.roots[ci], L, 0); launches++;
    st.W_log2n[ci] = L;
  }
  // input -> staging buffer (host input) ; bit reversal writes the working array
  const void* d_in = in;
  if (!is_dev(in)) { int rc = ensure(ctx, &st.buf, &st.buf_cap, bytes); if (rc) return rc; NCK(cudaMemcpyAsync(st.buf, in, bytes, cudaMemcpyHostToDevice, s)); d_in = st.buf; }
  void* d_x;
  if (is_dev(out) && out != in) d_x = out;
  else { int rc = ensure(ctx, &st.buf2, &st.buf2_cap, bytes); if (rc) return rc; d_x = st.buf2; }
  for (auto& e : st.ev) if (!e) NCK(cudaEventCreate(&e));
  NCK(cudaEventRecord(st.ev[0], s));
  const uint32_t K = L < (uint32_t)TILE_LOG ? L : (uint32_t)TILE_LOG;
  if (K == 0) { k_ntt_bitrev<F><<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(d_in, d_x, L); launches++; }      // n == 1: plain copy
  NCK(cudaEventRecord(st.ev[1], s));
  if (K >= 1) { k_ntt_tile<F><<<(uint32_t)(n >> K), TILE / 4, 0, s>>>(d_in, d_x, st.W[ci], L, K); launches++; }
  NCK(cudaEventRecord(st.ev[2], s));
  uint32_t sg = K + 1; st.last_passes4 = st.last_passes2 = 0;
  for (; sg + 1 <= L; sg += 2) { k_ntt_stage4<F><<<(uint32_t)((n / 4 + 255) / 256), 256, 0, s>>>(d_x, st.W[ci], L, sg); launches++; st.last_passes4++; }
  if (sg <= L) { k_ntt_stage2<F><<<(uint32_t)((n / 2 + 255) / 256), 256, 0, s>>>(d_x, st.W[ci], L, sg); launches++; st.last_passes2++; }
  NCK(cudaEventRecord(st.ev[3], s));
  if (inverse && L >= 1) { k_ntt_final_inverse<F><<<(uint32_t)((n / 2 + 1 + 255) / 256), 256, 0, s>>>(d_x, aux, L); launches++; }
  NCK(cudaEventRecord(st.ev[4], s));
  NCK(cudaGetLastError());
  b200msm_internal_count_launches(ctx, launches);
  if (d_x != out) {
    NCK(cudaMemcpyAsync(out, d_x, bytes, is_dev(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    if (!is_dev(out)) NCK(cudaStreamSynchronize(s));
  }
  return B200MSM_OK;
}

}  // namespace

void b200ntt_release(b200msm_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_state.find(ctx);
  if (it == g_state.end()) return;
  cudaSetDevice(b200msm_internal_device(ctx));
  for (int c = 0; c < 2; c++) { if (it->second.roots[c]) cudaFree(it->second.roots[c]); if (it->second.W[c]) cudaFree(it->second.W[c]); }
  if (it->second.buf) cudaFree(it->second.buf);
  if (it->second.buf2) cudaFree(it->second.buf2);
  for (auto& e : it->second.ev) if (e) cudaEventDestroy(e);
  g_state.erase(it);
}

extern "C" int b200msm_fr_fft(b200msm_ctx* ctx, int curve, const void* in, uint32_t log2n, int inverse, void* out) {
  if (!ctx || !in || !out || (curve != 0 && curve != 1)) return B200MSM_E_ARG;
  const uint32_t maxbits = curve == 0 ? BLS12_381_FR::MAXBITS : BN254_FR::MAXBITS;
  if (log2n > maxbits || log2n > 28) { b200msm_internal_set_error(ctx, "transform size exceeds the 2-adicity of Fr (or 2^28)"); return B200MSM_E_UNSUPPORTED; }
  if (cudaSetDevice(b200msm_internal_device(ctx)) != cudaSuccess) return B200MSM_E_CUDA;
  NttState* st;
  { std::lock_guard<std::mutex> lk(g_mu); st = &g_state[ctx]; }
  return curve == 0 ? run_ntt<BLS12_381_FR>(ctx, *st, 0, in, log2n, inverse, out) : run_ntt<BN254_FR>(ctx, *st, 1, in, log2n, inverse, out);
}

// phase times of the last transform on this context (milliseconds): [0] bit reversal, [1] shared-memory tile stages,
// [2] radix-4 / radix-2 global passes, [3] inverse finalisation; passes[0] = radix-4 passes, passes[1] = radix-2 passes
extern "C" int b200msm_fr_fft_last_phases(b200msm_ctx* ctx, float ms[4], uint32_t passes[2]) {
  if (!ctx || !ms || !passes) return B200MSM_E_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_state.find(ctx);
  if (it == g_state.end() || !it->second.ev[4]) return B200MSM_E_ARG;
  if (cudaEventSynchronize(it->second.ev[4]) != cudaSuccess) return B200MSM_E_CUDA;
  for (int k = 0; k < 4; k++) cudaEventElapsedTime(&ms[k], it->second.ev[k], it->second.ev[k + 1]);
  passes[0] = it->second.last_passes4; passes[1] = it->second.last_passes2;
  return B200MSM_OK;
}
