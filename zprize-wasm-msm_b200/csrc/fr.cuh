// fr.cuh -- the scalar fields Fr of BLS12-381 and BN254 as field classes for the templates of fp.cuh (8 u32 limbs, Montgomery R = 2^256),
// with the 2-adic root of unity the reference's FFT is built on: w[maxBits] = nr^((r-1)/2^maxBits), nr = smallest quadratic non-residue
// (wasmcurves/src/build_fft.js:32-52; frm is wired at src/bls12381/build_bls12381.js:39-43 and src/bn128/build_bn128.js:35-39).
#pragma once
#include "fp.cuh"

namespace b200 {

}  // namespace b200
