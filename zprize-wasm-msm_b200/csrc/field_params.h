// field_params.h -- the field constants of the path as constexpr classes, usable from host C++ (g++) and from CUDA.
//
// Base fields Fq of BLS12-381 / BN254 (wasmcurves/src/bls12381/build_bls12381.js:22, src/bn128/build_bn128.js:20) and the scalar
// fields Fr the reference's FFT runs over (build_bls12381.js:39-43, build_bn128.js:35-39).  Every class gives N u32 limbs
// (little-endian), NP = -q^-1 mod 2^32 (build_f1m.js:504), R mod q ("one"), R^2 mod q, R^3 mod q with R = 2^(32 N).
// This header has no device code: tests/host_inv_harness.cpp compiles it (with bingcd.h) under plain g++.
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__)
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace b200 {

struct BLS12_381 {
  static constexpr int ID = 0;
  static constexpr int EXT = 1;          // prime field
  static constexpr int N = 12;           // u32 limbs per Fq element (n8 = 48)
  static constexpr uint32_t NP = 0xfffcfffdu;   // -q^-1 mod 2^32   (build_f1m.js:504)
  static constexpr int QBITS = 381;
  __host__ __device__ static constexpr uint32_t q(int i) {      // build_bls12381.js:22
    constexpr uint32_t t[N] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                               0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return t[i];
  }
  __host__ __device__ static constexpr uint32_t one(int i) {    // R mod q
    constexpr uint32_t t[N] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                               0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
    return t[i];
  }
  __host__ __device__ static constexpr uint32_t r2(int i) {     // R^2 mod q
    constexpr uint32_t t[N] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                               0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
    return t[i];
  }
  __host__ __device__ static constexpr uint32_t r3(int i) {     // R^3 mod q
    constexpr uint32_t t[N] = {0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au,
                               0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u};
    return t[i];
  }
};

struct BN254 {
  static constexpr int ID = 1;
  static constexpr int EXT = 1;
  static constexpr int N = 8;            // n8 = 32
  static constexpr uint32_t NP = 0xe4866389u;
  static constexpr int QBITS = 254;
  __host__ __device__ static constexpr uint32_t q(int i) {      // build_bn128.js:20
    constexpr uint32_t t[N] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  __host__ __device__ static constexpr uint32_t one(int i) {
    constexpr uint32_t t[N] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  __host__ __device__ static constexpr uint32_t r2(int i) {
    constexpr uint32_t t[N] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return t[i];
  }
  __host__ __device__ static constexpr uint32_t r3(int i) {
    constexpr uint32_t t[N] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
    return t[i];
  }
};

// ---- scalar fields (fr.cuh adds nothing but the include): 2-adic roots of unity as in wasmcurves/src/build_fft.js:32-52
struct BLS12_381_FR {
  static constexpr int ID = 0, EXT = 1, N = 8, QBITS = 255;
  static constexpr uint32_t NP = 0xffffffffu;          // -r^-1 mod 2^32
  static constexpr int MAXBITS = 32;                    // r - 1 = 2^32 * odd
  __host__ __device__ static constexpr uint32_t q(int i) {
    constexpr uint32_t t[N] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t one(int i) {
    constexpr uint32_t t[N] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t r2(int i) {
    constexpr uint32_t t[N] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t r3(int i) {
    constexpr uint32_t t[N] = {0x439b73afu, 0xc62c1807u, 0x8cf06990u, 0x1b3e0d18u, 0xc7b5f418u, 0x73d13c71u, 0xc8db33e9u, 0x6e2a5bb9u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t root(int i) {      // 5^((r-1)/2^32) * R mod r: primitive 2^32-th root of unity
    constexpr uint32_t t[N] = {0x0c17f47cu, 0x9cab6d5cu, 0xfd4b71e5u, 0x1ce1e93du, 0x471dd505u, 0x0d6db230u, 0x743a3b6au, 0x3f0ee990u}; return t[i]; }
};

struct BN254_FR {
  static constexpr int ID = 1, EXT = 1, N = 8, QBITS = 254;
  static constexpr uint32_t NP = 0xefffffffu;
  static constexpr int MAXBITS = 28;
  __host__ __device__ static constexpr uint32_t q(int i) {
    constexpr uint32_t t[N] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t one(int i) {
    constexpr uint32_t t[N] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t r2(int i) {
    constexpr uint32_t t[N] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}; return t[i]; }
  __host__ __device__ static constexpr uint32_t r3(int i) {
    constexpr uint32_t t[N] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu}; return t[i]; }
  __host__ __device__ static constexpr uint32_t root(int i) {      // 5^((r-1)/2^28) * R mod r
    constexpr uint32_t t[N] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u}; return t[i]; }
};

}  // namespace b200
