// probes.cu -- does the FP64 pipe run beside the integer-multiply pipe?  
//
// The batch-affine kernels are bound by IMAD.WIDE issue (fp.cuh: one BLS12-381 Fq multiplication = 303 multiplier
// instructions, 99 % of the pipe).  B200 also has a full-rate FP64 pipe that the integer path leaves idle.  A Montgomery
// multiplication on 8 x 48-bit limbs held in doubles costs about 4 FP64 operations per limb product (fma.rz against
// 2^100 to accumulate the high halves, one subtraction to isolate the high half, one fma for the low half, one add),
// ~136 limb products + per-row reduction ~ 690 FP64 operations.  Before building that multiplier this probe measures
// the only thing that decides whether it pays: the time of
//   (a) `iters` real Fq multiplications per thread (IMAD.WIDE carry chains),
//   (b) `iters` blocks of 690 dependent-chain FP64 operations per thread (the instruction mix of the FP64 multiplier:
//       2/3 DFMA, 1/3 DADD over 32 accumulators),
//   (c) both in the same thread, the compiler free to interleave them.
// If (c) ~ max(a, b) the pipes overlap and routing part of the multiplications through FP64 buys throughput;
// if (c) ~ a + b they share an issue bottleneck and the idea is dead.
#include <cuda_runtime.h>
#include <stdint.h>
#include "internal.h"
#include "../../include/b200msm_probes.h"
#include "fp.cuh"

using namespace b200;

namespace {

constexpr int F64_ACC = 32;            // column accumulators of the FP64 multiplier (16 high + 16 low)
constexpr int F64_OPS = 690;           // FP64 operations per multiplication (see above)
static_assert(21 * F64_ACC + 18 == F64_OPS, "the unrolled block below issues F64_OPS operations");

template <int MODE>   // 1: integer only, 2: FP64 only, 3: both in every thread, 4: odd warps integer, even warps FP64 (warp-specialised)
__global__ void __launch_bounds__(256) k_dualpipe(uint32_t iters, const void* __restrict__ in, void* __restrict__ out, double seed) {
  using C = BLS12_381;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe<C::N> x, y;
  fe_load<C>(x, reinterpret_cast<const char*>(in) + (uint64_t)(i & 1023) * 4 * C::N);
  y = x;
  double acc[F64_ACC], a = seed + (double)(i & 7), b = seed * 1.000001;
#pragma unroll
  for (int k = 0; k < F64_ACC; k++) acc[k] = seed + k;
  const bool do_int = MODE == 4 ? ((threadIdx.x >> 5) & 1) : (MODE & 1), do_f64 = MODE == 4 ? !((threadIdx.x >> 5) & 1) : (MODE & 2);
  for (uint32_t it = 0; it < iters; it++) {
    if (do_int) fe_mul<C>(y, y, x);
    if (do_f64) {
      // 690 operations: 21 sweeps over the 32 accumulators (alternating fma, fma, add) + 18 more
#pragma unroll
      for (int r = 0; r < 21; r++) {
#pragma unroll
        for (int k = 0; k < F64_ACC; k++) {
          if ((r + k) % 3 == 2) acc[k] = __dadd_rn(acc[k], a);
          else acc[k] = __fma_rz(a, b, acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 18; k++) acc[k] = __fma_rz(a, b, acc[k]);
      a = acc[it & 1] * 1e-300 + seed;       // keeps the chain dependent on the results without letting values explode
    }
  }
  double t = 0;
#pragma unroll
  for (int k = 0; k < F64_ACC; k++) t += acc[k];
  if (t == 12345.6789) y.l[0] ^= 1;
  fe_store<C>(reinterpret_cast<char*>(out) + (uint64_t)i * 4 * C::N, y);
}

}  // namespace

// out[0] = integer-only ms, out[1] = FP64-only ms, out[2] = both in one thread ms, out[3] = multiplications per thread,
// out[4] = warp-specialised ms (half of the warps do the integer work, the other half the FP64 work: half of each pipe's load)
extern "C" int b200msm_probe_dualpipe(b200msm_ctx* ctx, double out[5]) {
  if (!ctx || !out) return B200MSM_E_ARG;
  if (cudaSetDevice(b200msm_internal_device(ctx)) != cudaSuccess) return B200MSM_E_CUDA;
  cudaStream_t s = b200msm_internal_stream(ctx);
  cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, b200msm_internal_device(ctx)) != cudaSuccess) return B200MSM_E_CUDA;
  const uint32_t blocks = prop.multiProcessorCount * 4, threads = 256, iters = 256;
  void *din = nullptr, *dout = nullptr;
  if (cudaMalloc(&din, 1024 * 48) != cudaSuccess || cudaMalloc(&dout, (size_t)blocks * threads * 48) != cudaSuccess) { cudaFree(din); return B200MSM_E_NOMEM; }
  cudaMemsetAsync(din, 0x17, 1024 * 48, s);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 1; mode <= 4; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, s);
      if (mode == 1) k_dualpipe<1><<<blocks, threads, 0, s>>>(iters, din, dout, 1.000001);
      else if (mode == 2) k_dualpipe<2><<<blocks, threads, 0, s>>>(iters, din, dout, 1.000001);
      else if (mode == 3) k_dualpipe<3><<<blocks, threads, 0, s>>>(iters, din, dout, 1.000001);
      else k_dualpipe<4><<<blocks, threads, 0, s>>>(iters, din, dout, 1.000001);
      cudaEventRecord(e1, s); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    out[mode == 4 ? 4 : mode - 1] = best;
  }
  out[3] = iters;
  b200msm_internal_count_launches(ctx, 12);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(din); cudaFree(dout);
  return cudaGetLastError() == cudaSuccess ? B200MSM_OK : B200MSM_E_CUDA;
}
