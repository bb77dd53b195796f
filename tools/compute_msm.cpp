// compute_msm.cpp -- a compiled C++ host for the B200 MSM engine: the ZPrize-style `compute_msm(bases, scalars)` entry point over the C ABI.
//
// The reference repository's own native host is a placeholder (/root/reference/src/main.rs:1-3); its working hosts are the JavaScript rigs
// (wasmcurves/benchmarks/multiexp.js:7-42, test/batchAffine.js:1177-1255) that hand pBases / pScalars to g1m_multiexpAffine and read one point back.
// This program is that caller in C++: it binds include/b200msm.h only (no Python, no torch), so it is also the smallest example of what a
// JS (N-API) / Rust (FFI) host links against.
//
//   compute_msm <bls12381|bn128> <log2n> <seed> [--devices 0,1,...] [--bases-on-host] [--repeat K]
//
// Inputs: bases P_i = splitmix64(seed + i) * G from the engine's generator (the same stream as oracle/msm_oracle.c), scalars = 4 splitmix64
// words each from the stream seed ^ 0x5ca1ab1e.  Output (stdout): one line `x=<hex> y=<hex>` -- the canonical affine result, plain integers --
// and one line `ms=<best milliseconds per MSM>`.  Exit status 2 when no usable GPU exists (there is no CPU fallback).
//
// Build: g++ -O2 -std=c++17 tools/compute_msm.cpp -Iinclude -I/usr/local/cuda/include -L<dir of libb200msm.so> -lb200msm -L/usr/local/cuda/lib64 -lcudart -o compute_msm
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime_api.h>
#include "b200msm.h"

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// == the harness entry point: sum_i scalars[i] * bases[i] -> canonical affine (x, y), 2*n8 bytes.  bases: n affine Montgomery points (host or device
// memory), scalars: n x 32 bytes little-endian (host or device).  Returns a b200msm status.
static int compute_msm(b200msm_ctx* ctx, int curve, const void* bases, const void* scalars, uint64_t n, void* xy_out) {
  uint8_t jac[3 * 96];
  int rc = b200msm_g1_multiexp_affine(ctx, curve, bases, scalars, 32, n, jac);
  if (rc) return rc;
  return b200msm_g1_normalize(ctx, curve, jac, 1, xy_out);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <bls12381|bn128> <log2n> <seed> [--devices 0,1,...] [--bases-on-host] [--repeat K]\n", argv[0]); return 1; }
  const int curve = !strcmp(argv[1], "bn128") ? B200MSM_BN254_G1 : B200MSM_BLS12_381_G1;
  const int log2n = atoi(argv[2]); const uint64_t seed = strtoull(argv[3], nullptr, 0);
  if (log2n < 0 || log2n > 28) { fprintf(stderr, "log2n out of range\n"); return 1; }
  std::vector<int> devs; bool host_bases = false; int repeat = 1;
  for (int i = 4; i < argc; i++) {
    if (!strcmp(argv[i], "--devices") && i + 1 < argc) { for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devs.push_back(atoi(t)); }
    else if (!strcmp(argv[i], "--bases-on-host")) host_bases = true;
    else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
  }
  if (devs.empty()) devs.push_back(0);
  const uint64_t n = 1ull << log2n; const size_t n8 = curve == B200MSM_BN254_G1 ? 32 : 48;

  b200msm_ctx* ctx = nullptr;
  int rc = b200msm_create_multi(&ctx, devs.data(), (int)devs.size());
  if (rc) { fprintf(stderr, "compute_msm: %s -- no usable GPU, and there is no CPU fallback\n", b200msm_strerror(rc)); return 2; }

  // inputs
  void* d_bases = nullptr;
  if (cudaSetDevice(devs[0]) != cudaSuccess || cudaMalloc(&d_bases, n * 2 * n8 + 16) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 2; }
  rc = b200msm_g1_generate_bases(ctx, curve, seed, 0, n, d_bases);
  if (rc) { fprintf(stderr, "generate_bases: %s (%s)\n", b200msm_strerror(rc), b200msm_last_error(ctx)); return 3; }
  std::vector<uint64_t> scalars(n * 4);
  for (uint64_t i = 0; i < n * 4; i++) scalars[i] = splitmix64((seed ^ 0x5ca1ab1eull) + i);
  std::vector<uint8_t> h_bases;
  const void* bases = d_bases;
  if (host_bases || devs.size() > 1) {      // several GPUs: each pulls its slice of the HOST buffers over its own PCIe link
    h_bases.resize(n * 2 * n8);
    if (cudaMemcpy(h_bases.data(), d_bases, n * 2 * n8, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "cudaMemcpy failed\n"); return 3; }
    bases = h_bases.data();
  }

  std::vector<uint8_t> xy(2 * n8);
  double best = 1e30;
  for (int k = 0; k < repeat; k++) {
    auto t0 = std::chrono::steady_clock::now();
    rc = compute_msm(ctx, curve, bases, scalars.data(), n, xy.data());
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc) { fprintf(stderr, "compute_msm: %s (%s)\n", b200msm_strerror(rc), b200msm_last_error(ctx)); return 3; }
    if (ms < best) best = ms;
  }
  auto hex = [&](const uint8_t* p) { std::string s; char b[3]; for (int i = (int)n8 - 1; i >= 0; i--) { snprintf(b, sizeof b, "%02x", p[i]); s += b; } return s; };
  printf("x=%s y=%s\n", hex(xy.data()).c_str(), hex(xy.data() + n8).c_str());
  printf("ms=%.3f devices=%d n=%llu\n", best, b200msm_device_count(ctx), (unsigned long long)n);
  cudaFree(d_bases);
  b200msm_destroy(ctx);
  return 0;
}
