// host_inv_harness.cpp -- CPU harness for zprize-wasm-msm_b200/csrc/bingcd.h (the root inversion of every batch inversion).
// Built by tests/test_host_inv.py with g++; the SAME header is compiled by nvcc into the engine (fe_inv_fast in fp.cuh).
//   field: 0 BLS12-381 Fq, 1 BN254 Fq, 2 BLS12-381 Fr, 3 BN254 Fr.  in / out: count residues of 4*N bytes, plain (not Montgomery).
#include <stdint.h>
#include <string.h>
#include "field_params.h"
#include "bingcd.h"

using namespace b200;

template <class C> static void run(const uint8_t* in, uint8_t* out, uint64_t count) {
  for (uint64_t k = 0; k < count; k++) {
    uint32_t y[C::N], r[C::N];
    memcpy(y, in + k * 4 * C::N, 4 * C::N);
    bingcd_inverse<C>(r, y);
    memcpy(out + k * 4 * C::N, r, 4 * C::N);
  }
}
extern "C" int host_inverse(int field, const uint8_t* in, uint8_t* out, uint64_t count) {
  switch (field) {
    case 0: run<BLS12_381>(in, out, count); return 0;
    case 1: run<BN254>(in, out, count); return 0;
    case 2: run<BLS12_381_FR>(in, out, count); return 0;
    case 3: run<BN254_FR>(in, out, count); return 0;
  }
  return -1;
}
extern "C" int host_field_info(int field, uint32_t* n, uint8_t* q) {
  switch (field) {
    case 0: *n = BLS12_381::N; for (int i = 0; i < BLS12_381::N; i++) { uint32_t v = BLS12_381::q(i); memcpy(q + 4 * i, &v, 4); } return 0;
    case 1: *n = BN254::N; for (int i = 0; i < BN254::N; i++) { uint32_t v = BN254::q(i); memcpy(q + 4 * i, &v, 4); } return 0;
    case 2: *n = BLS12_381_FR::N; for (int i = 0; i < BLS12_381_FR::N; i++) { uint32_t v = BLS12_381_FR::q(i); memcpy(q + 4 * i, &v, 4); } return 0;
    case 3: *n = BN254_FR::N; for (int i = 0; i < BN254_FR::N; i++) { uint32_t v = BN254_FR::q(i); memcpy(q + 4 * i, &v, 4); } return 0;
  }
  return -1;
}
