/* b200msm_napi.c -- thin N-API addon: the JavaScript-facing shim over the C ABI (include/b200msm.h).
 *
 * Exposes to Node.js the calls a wasmcurves / ffjavascript host makes on the MSM path, with Buffers instead of
 * linear-memory pointers (wasmcurves/src/build_multiexp.js:251-371, :96-249):
 *
 *   const msm = require("./b200msm.node");
 *   const ctx = msm.create(0);                                      // one GPU, like one WASM instance
 *   const pr  = msm.g1m_multiexpAffine(ctx, "bls12381", bases, scalars, 32, n);          // Buffer(3*n8), Jacobian Montgomery
 *   const pc  = msm.g1m_multiexpAffine_chunk(ctx, "bls12381", bases, scalars, 32, n, startBit, chunkSize);
 *   const xy  = msm.g1m_normalize(ctx, "bls12381", pr);                                  // Buffer(2*n8), canonical affine
 *
 * Node.js and node_api.h are not present in the build image, so this file is compiled only where they are
 * (cc -shared -fPIC -I$(node -p "require('node-api-headers').include_dir") b200msm_napi.c -L. -lb200msm -o b200msm.node);
 * every function is a mechanical marshalling wrapper, no arithmetic happens here.
 */
#if defined(B200MSM_BUILD_NAPI) || defined(NAPI_VERSION)
#include <node_api.h>
#include <string.h>
#include <stdint.h>
#include "../../include/b200msm.h"

#define NAPI_CALL(env, call) do { if ((call) != napi_ok) { napi_throw_error((env), NULL, "N-API call failed: " #call); return NULL; } } while (0)

static int curve_of(napi_env env, napi_value v) {
  char buf[16]; size_t len = 0;
  if (napi_get_value_string_utf8(env, v, buf, sizeof buf, &len) != napi_ok) return -1;
  if (!strcmp(buf, "bls12381")) return B200MSM_BLS12_381_G1;
  if (!strcmp(buf, "bn128") || !strcmp(buf, "bn254")) return B200MSM_BN254_G1;
  return -1;
}
static napi_value throw_status(napi_env env, b200msm_ctx* ctx, int rc) {
  char msg[512]; strncpy(msg, b200msm_strerror(rc), sizeof msg - 1); msg[sizeof msg - 1] = 0;
  if (ctx) { strncat(msg, ": ", sizeof msg - strlen(msg) - 1); strncat(msg, b200msm_last_error(ctx), sizeof msg - strlen(msg) - 1); }
  napi_throw_error(env, NULL, msg); return NULL;
}
static void ctx_finalize(napi_env env, void* data, void* hint) { (void)env; (void)hint; b200msm_destroy((b200msm_ctx*)data); }

static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; int32_t dev = -1;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc >= 1) napi_get_value_int32(env, argv[0], &dev);
  b200msm_ctx* ctx = NULL; int rc = b200msm_create(&ctx, dev);
  if (rc) return throw_status(env, NULL, rc);
  napi_value ext; NAPI_CALL(env, napi_create_external(env, ctx, ctx_finalize, NULL, &ext));
  return ext;
}

/* shared body of g1m_/g2m_ multiexpAffine[_chunk] and multiexp[_chunk]: chunk = per-window form, g2 = G2 exports (Fq2 elements),
 * jac = Jacobian bases (3 elements per point instead of 2) */
static napi_value Multiexp(napi_env env, napi_callback_info info, int chunk, int g2, int jac) {
  size_t argc = 8; napi_value argv[8];
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < (size_t)(chunk ? 8 : 6)) { napi_throw_type_error(env, NULL, "too few arguments"); return NULL; }
  b200msm_ctx* ctx; NAPI_CALL(env, napi_get_value_external(env, argv[0], (void**)&ctx));
  int curve = curve_of(env, argv[1]); if (curve < 0) { napi_throw_type_error(env, NULL, "curve must be 'bls12381' or 'bn128'"); return NULL; }
  void *bases, *scalars; size_t nb, ns; uint32_t ssz, start = 0, bits = 0; int64_t n;
  NAPI_CALL(env, napi_get_buffer_info(env, argv[2], &bases, &nb));
  NAPI_CALL(env, napi_get_buffer_info(env, argv[3], &scalars, &ns));
  NAPI_CALL(env, napi_get_value_uint32(env, argv[4], &ssz));
  NAPI_CALL(env, napi_get_value_int64(env, argv[5], &n));
  if (chunk) { NAPI_CALL(env, napi_get_value_uint32(env, argv[6], &start)); NAPI_CALL(env, napi_get_value_uint32(env, argv[7], &bits)); }
  const size_t n8 = (curve == B200MSM_BLS12_381_G1 ? 48 : 32) * (g2 ? 2 : 1);      /* bytes per coordinate-field element */
  const int cid = curve + (g2 ? 2 : 0);                                            /* B200MSM_*_G2 = G1 id + 2 */
  if (n < 0 || nb < (size_t)n * (jac ? 3 : 2) * n8 || ns < (size_t)n * ssz) { napi_throw_range_error(env, NULL, "buffers shorter than n points / scalars"); return NULL; }
  void* out; napi_value res; NAPI_CALL(env, napi_create_buffer(env, 3 * n8, &out, &res));
  int rc;
  if (jac) rc = chunk ? b200msm_g1_multiexp_chunk(ctx, cid, bases, scalars, ssz, (uint64_t)n, start, bits, out)
                      : b200msm_g1_multiexp(ctx, cid, bases, scalars, ssz, (uint64_t)n, out);
  else rc = chunk ? b200msm_g1_multiexp_affine_chunk(ctx, cid, bases, scalars, ssz, (uint64_t)n, start, bits, out)
                  : b200msm_g1_multiexp_affine(ctx, cid, bases, scalars, ssz, (uint64_t)n, out);
  if (rc) return throw_status(env, ctx, rc);
  return res;
}
static napi_value MultiexpAffine(napi_env env, napi_callback_info info) { return Multiexp(env, info, 0, 0, 0); }
static napi_value MultiexpAffineChunk(napi_env env, napi_callback_info info) { return Multiexp(env, info, 1, 0, 0); }
static napi_value MultiexpJac(napi_env env, napi_callback_info info) { return Multiexp(env, info, 0, 0, 1); }
static napi_value MultiexpJacChunk(napi_env env, napi_callback_info info) { return Multiexp(env, info, 1, 0, 1); }
static napi_value G2MultiexpAffine(napi_env env, napi_callback_info info) { return Multiexp(env, info, 0, 1, 0); }
static napi_value G2MultiexpAffineChunk(napi_env env, napi_callback_info info) { return Multiexp(env, info, 1, 1, 0); }
static napi_value G2MultiexpJac(napi_env env, napi_callback_info info) { return Multiexp(env, info, 0, 1, 1); }

/* frm_fft(ctx, curve, buffer) / frm_ifft: n = buffer length / 32 must be a power of two; returns a new Buffer */
static napi_value FrFft(napi_env env, napi_callback_info info, int inverse) {
  size_t argc = 3; napi_value argv[3];
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 3) { napi_throw_type_error(env, NULL, "too few arguments"); return NULL; }
  b200msm_ctx* ctx; NAPI_CALL(env, napi_get_value_external(env, argv[0], (void**)&ctx));
  int curve = curve_of(env, argv[1]); if (curve < 0) { napi_throw_type_error(env, NULL, "bad curve"); return NULL; }
  void* in; size_t len; NAPI_CALL(env, napi_get_buffer_info(env, argv[2], &in, &len));
  size_t n = len / 32; uint32_t lg = 0; while (((size_t)1 << lg) < n) lg++;
  if (n == 0 || len != n * 32 || ((size_t)1 << lg) != n) { napi_throw_range_error(env, NULL, "buffer must hold a power-of-two number of 32-byte elements"); return NULL; }
  void* out; napi_value res; NAPI_CALL(env, napi_create_buffer(env, len, &out, &res));
  int rc = b200msm_fr_fft(ctx, curve, in, lg, inverse, out);
  if (rc) return throw_status(env, ctx, rc);
  return res;
}
static napi_value FrmFft(napi_env env, napi_callback_info info) { return FrFft(env, info, 0); }
static napi_value FrmIfft(napi_env env, napi_callback_info info) { return FrFft(env, info, 1); }

static napi_value Normalize(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  b200msm_ctx* ctx; NAPI_CALL(env, napi_get_value_external(env, argv[0], (void**)&ctx));
  int curve = curve_of(env, argv[1]); if (curve < 0) { napi_throw_type_error(env, NULL, "bad curve"); return NULL; }
  void* jac; size_t len; NAPI_CALL(env, napi_get_buffer_info(env, argv[2], &jac, &len));
  const size_t n8 = curve == B200MSM_BLS12_381_G1 ? 48 : 32;
  if (len % (3 * n8)) { napi_throw_range_error(env, NULL, "buffer is not a whole number of Jacobian points"); return NULL; }
  void* out; napi_value res; NAPI_CALL(env, napi_create_buffer(env, len / 3 * 2, &out, &res));
  int rc = b200msm_g1_normalize(ctx, curve, jac, len / (3 * n8), out);
  if (rc) return throw_status(env, ctx, rc);
  return res;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
    {"create", NULL, Create, NULL, NULL, NULL, napi_default, NULL},
    {"g1m_multiexpAffine", NULL, MultiexpAffine, NULL, NULL, NULL, napi_default, NULL},
    {"g1m_multiexpAffine_chunk", NULL, MultiexpAffineChunk, NULL, NULL, NULL, napi_default, NULL},
    {"g1m_normalize", NULL, Normalize, NULL, NULL, NULL, napi_default, NULL},
    {"g1m_multiexp", NULL, MultiexpJac, NULL, NULL, NULL, napi_default, NULL},
    {"g1m_multiexp_chunk", NULL, MultiexpJacChunk, NULL, NULL, NULL, napi_default, NULL},
    {"g2m_multiexpAffine", NULL, G2MultiexpAffine, NULL, NULL, NULL, napi_default, NULL},
    {"g2m_multiexpAffine_chunk", NULL, G2MultiexpAffineChunk, NULL, NULL, NULL, napi_default, NULL},
    {"g2m_multiexp", NULL, G2MultiexpJac, NULL, NULL, NULL, napi_default, NULL},
    {"frm_fft", NULL, FrmFft, NULL, NULL, NULL, napi_default, NULL},
    {"frm_ifft", NULL, FrmIfft, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
#endif /* B200MSM_BUILD_NAPI */
