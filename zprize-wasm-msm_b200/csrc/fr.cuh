// fr.cuh -- the scalar fields Fr of BLS12-381 and BN254 as field classes for the templates of fp.cuh (8 u32 limbs, Montgomery R = 2^256),
// with the 2-adic root of unity the reference's FFT is built on: w[maxBits] = nr^((r-1)/2^maxBits), nr = smallest quadratic non-residue
// (wasmcurves/src/build_fft.js:32-52; frm is wired at src/bls12381/build_bls12381.js:39-43 and src/bn128/build_bn128.js:35-39).
#pragma once
#include "fp.cuh"

namespace b200 {

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
