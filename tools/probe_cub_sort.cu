// probe_cub_sort.cu -- how fast is the "named" sort on this GPU?  (development probe, standalone executable)
//
// north_star (2) names a radix sort of (bucket, point) pairs with warp-level primitives and shared-memory staging.  The best available
// implementation of exactly that is CUB's onesweep DeviceRadixSort; this probe times it on the engine's own problem shape -- n_pairs
// (bucket id, point index) pairs with `bits`-bit keys (2^20 points x 16 windows, 17 x 2^15 buckets -> 20-bit keys) -- so that the
// engine's atomic-rank sort (k_digits<count> + scan + k_digits<scatter>) is compared with a measured number, not an estimate.
// The sort core alone is timed: a complete replacement would add a digit pass that writes the 8-byte pairs (134 MB at 2^20) and a pass
// that finds the bucket boundaries in the sorted keys.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/probe_cub_sort tools/probe_cub_sort.cu
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
__global__ void k_fill(uint32_t* keys, uint32_t* vals, uint64_t n, uint32_t nbuckets) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t x = i * 0x9E3779B97F4A7C15ull + 12345; x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
  keys[i] = (uint32_t)(x % nbuckets); vals[i] = (uint32_t)i;
}
int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : (16ull << 20);
  const uint32_t nbuckets = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 10) : 17u << 15;
  int bits = 0; while ((1ull << bits) < nbuckets) bits++;
  uint32_t *k0, *k1, *v0, *v1; void* tmp = nullptr; size_t tb = 0;
  cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
  cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, (int)n, 0, bits);
  cudaMalloc(&tmp, tb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 8; rep++) {
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(k0, v0, n, nbuckets);
    cudaEventRecord(e0);
    cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, (int)n, 0, bits);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep >= 2 && ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  printf("{\"probe\": \"cub_radix_sort_pairs\", \"pairs\": %llu, \"buckets\": %u, \"key_bits\": %d, \"ms\": %.4f, \"pairs_per_s\": %.3e, \"temp_bytes\": %zu, \"cuda\": \"%s\"}\n",
         (unsigned long long)n, nbuckets, bits, best, n / (best * 1e-3), tb, cudaGetErrorString(err));
  return 0;
}
