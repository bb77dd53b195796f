/* b200msm.h -- C ABI of the B200-native G1 multi-scalar-multiplication engine.
 *
 * Drop-in boundary for the MSM hot path of Manta-Network/zprize-wasm-msm (wasmcurves fork).
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/wasmcurves/).  The reference's "plugin API" is the WASM export table plus one
 * linear memory; here pointers are 64-bit, every call returns a status, and an
 * explicit context owns the GPU stream and scratch memory.
 *
 * Data formats are the reference's, byte for byte:
 *   Fq element  : n8 bytes little-endian, Montgomery form a*R mod q, R = 2^(8*n8)   (src/build_f1m.js:30-43)
 *                 n8 = 48 for BLS12-381, 32 for BN254 ("bn128")
 *   affine point: x || y (2*n8 bytes); infinity = all zero                             (src/build_curve_jacobian_a0.js:55-77)
 *   Jacobian    : x || y || z (3*n8 bytes), (X/Z^2, Y/Z^3); infinity: z == 0, written as (0, R mod q, 0)  (:124-150)
 *   scalar      : scalar_size bytes (any size >= 1, as in the reference), plain unsigned little-endian integer, NOT Montgomery,
 *                 NOT required < r.  More than 32 bytes run as 256-bit slices combined by Horner.   (src/build_multiexp.js:73-92)
 * All `const void*` inputs and `void*` outputs may be HOST or DEVICE pointers (detected with
 * cudaPointerGetAttributes); device buffers must be 16-byte aligned.  There is no CPU fallback:
 * every compute entry point fails with B200MSM_E_CUDA when no usable GPU is present.
 * A context is not re-entrant (like a WASM instance, src/build_multiexp.js:273); use one per host thread.
 */
#ifndef B200MSM_H
#define B200MSM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200msm_ctx b200msm_ctx;

enum { B200MSM_BLS12_381_G1 = 0,   /* src/bls12381/build_bls12381.js:16-125 */
       B200MSM_BN254_G1 = 1,       /* src/bn128/build_bn128.js:14-125       */
       /* G2: the same entry points over the quadratic extension Fq2 = Fq[u]/(u^2+1) (src/build_f2m.js; g2m_* exports,
        * build_bls12381.js:48-53, build_bn128.js:44-49).  For these ids "n8" below is the size of an Fq2 element, c0 || c1 = 96 / 64
        * bytes, i.e. an affine point is 192 / 128 bytes and a Jacobian point 288 / 192 bytes -- the g2m layouts.  The MSM entry points
        * (== g2m_multiexpAffine, g2m_multiexpAffine_chunk), resident / windowed / batched forms, normalize, sum, generate_bases and
        * fq_op (== f2m_*) accept them; the point codecs and GLV are G1 only. */
       B200MSM_BLS12_381_G2 = 2,
       B200MSM_BN254_G2 = 3 };

enum { B200MSM_OK = 0, B200MSM_E_ARG = -1, B200MSM_E_CUDA = -2, B200MSM_E_NOMEM = -3, B200MSM_E_UNSUPPORTED = -4 };

/* Per-call measurements (milliseconds, CUDA events on the context's stream) and plan parameters. */
typedef struct b200msm_stats {
  uint32_t n, window_bits, windows, buckets_per_window, tree_rounds, reserved;
  uint64_t pairs;            /* non-zero (point, window) digits = bucket insertions                     */
  uint64_t affine_adds;      /* batch-affine additions executed                                          */
  float ms_total, ms_h2d, ms_digits_sort, ms_accumulate, ms_bucket_reduce, ms_window_combine, ms_d2h;
  /* kernel-group times inside the phases above (sums over all rounds / launches of that kernel group) */
  float ms_k_sort, ms_k_plan, ms_k_tree_fwd, ms_k_inv_tree, ms_k_tree_bwd, ms_k_finish, ms_k_fold, ms_k_wsum, ms_k_horner;
  float ms_host_combine;     /* host part of the window combination (inside ms_window_combine) */
  uint64_t launches;         /* kernels + memset/memcpy nodes launched by this call */
  uint64_t affine_adds_round0;   /* additions done by the first (largest) k_tree_bwd launch */
  float ms_k_tree_bwd_round0, reserved3;
} b200msm_stats;

/* ---- life cycle.  device_id < 0 selects the current CUDA device. */
int  b200msm_create(b200msm_ctx** out, int device_id);
/* One context over n_devices GPUs of this node (device_ids[0] is the "home" device that holds device-resident
 * outputs).  It owns one stream set, scratch pool and host thread per device.  Every MSM entry point below, called on such a context,
 * (ordinals may repeat: each entry is one shard with its own context) shards the points by contiguous range over the devices -- the counterpart of ffjavascript running g1m_multiexpAffine on a slice of
 * the inputs in each worker and adding the partial results (src/build_multiexp.js:319-369 is the per-worker part):
 * device g pulls ITS slice of the caller's (host) bases and scalars over its own PCIe link, runs the single-GPU pipeline, and the
 * n_devices partial points are added on the host.  Nothing but those 3*n8-byte partials is exchanged: there is no collective on the
 * data path.  b200msm_upload_bases[_windowed] on such a context shards the resident set the same way (option "multi_replicate" = 1
 * keeps a full copy on every device instead, which lets b200msm_g1_multiexp_batch run MSM j on device j mod n_devices).
 * Problems smaller than "multi_min_points" (default 2^15) per device use fewer devices.  Results are the same group element as on
 * one device.  n_devices == 1 is b200msm_create. */
int  b200msm_create_multi(b200msm_ctx** out, const int* device_ids, int n_devices);
int  b200msm_device_count(const b200msm_ctx* ctx);            /* devices behind this context (1 for b200msm_create) */
void b200msm_destroy(b200msm_ctx* ctx);
const char* b200msm_strerror(int status);
const char* b200msm_last_error(const b200msm_ctx* ctx);      /* detail string of the last failure */
const char* b200msm_version(void);
/* Run on an externally owned CUDA stream (cudaStream_t passed as void*), e.g. torch's current stream.  Work already issued on the
 * previous stream is drained first.  Multi-device contexts own their streams: B200MSM_E_UNSUPPORTED. */
int  b200msm_set_stream(b200msm_ctx* ctx, void* cuda_stream);
int  b200msm_synchronize(b200msm_ctx* ctx);

/* ---- == g1m_multiexpAffine(pBases, pScalars, scalarSize, n, pr)          src/build_multiexp.js:251-371
 *      == g1m_multiexp_multiExp / g1m_multiexpAffine_multiExp(pPoints, pScalars, numPoints, pResult)
 *                                                                         src/build_multiexp_opt.js:1987-2110 (scalar_size = 32)
 * out receives 3*n8 bytes: sum_i scalars[i] * bases[i] as a Jacobian Montgomery point; n == 0 -> canonical zero. */
int b200msm_g1_multiexp_affine(b200msm_ctx* ctx, int curve, const void* bases, const void* scalars,
                               uint32_t scalar_size, uint64_t n, void* out);

/* ---- == g1m_multiexpAffine_chunk(pBases, pScalars, scalarSize, n, startBit, chunkSize, pr)
 *                                                                         src/build_multiexp.js:96-249
 * One window: sum_i digit_i * bases[i], digit_i = bits [start_bit, start_bit + chunk_bits) of scalar i clipped at
 * the scalar end (src/build_multiexp.js:25-94), WITHOUT the 2^start_bit factor.  chunk_bits in [1, 32]. */
int b200msm_g1_multiexp_affine_chunk(b200msm_ctx* ctx, int curve, const void* bases, const void* scalars,
                                     uint32_t scalar_size, uint64_t n, uint32_t start_bit, uint32_t chunk_bits, void* out);

/* ---- == g1m_multiexp(pBases, pScalars, scalarSize, n, pr) / g1m_multiexp_chunk(..., startBit, chunkSize, pr): the same two functions
 *      over JACOBIAN bases of 3*n8 bytes each (n8b = 3*n8: src/build_curve_jacobian_a0.js:1429, src/build_multiexp.js:8-16); bases with
 *      z == 0 are the point at infinity and contribute nothing.  Also for the G2 curve ids (== g2m_multiexp). */
int b200msm_g1_multiexp(b200msm_ctx* ctx, int curve, const void* bases_jacobian, const void* scalars,
                        uint32_t scalar_size, uint64_t n, void* out);
int b200msm_g1_multiexp_chunk(b200msm_ctx* ctx, int curve, const void* bases_jacobian, const void* scalars,
                              uint32_t scalar_size, uint64_t n, uint32_t start_bit, uint32_t chunk_bits, void* out);

/* ---- resident bases: upload once (pb.alloc + pb.set of pBases in the reference rig, benchmarks/multiexp.js:16-23),
 *      then run any number of MSMs against them.  n in the MSM call may be <= the uploaded count. */
int b200msm_upload_bases(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, uint64_t* handle);
int b200msm_free_bases(b200msm_ctx* ctx, uint64_t handle);
int b200msm_g1_multiexp_resident(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size,
                                 uint64_t n, void* out, b200msm_stats* stats /* nullable */);

/* ---- resident bases WITH a precomputed window table (no counterpart in the reference, which receives its bases with
 *      every call; fixed-base provers upload once).  Row w of the table holds 2^(bit offset of window w) * P_i, so the digits
 *      of all windows share one bucket array: no per-window bucket reduction, no window combination, wider windows.
 *      The table serves MSMs whose scalar_size equals the one given here (other sizes run the ordinary pipeline on the
 *      same handle); window_bits = 0 chooses the width from n.  Device memory: ceil(8*scalar_size / window_bits) * n * 2*n8 bytes.
 *      Results are identical to b200msm_upload_bases + b200msm_g1_multiexp_resident. */
int b200msm_upload_bases_windowed(b200msm_ctx* ctx, int curve, const void* bases, uint64_t n, uint32_t scalar_size,
                                  uint32_t window_bits, uint64_t* handle);

/* ---- count independent MSMs of n points each over the same resident bases (a stream of proofs over one proving key; the
 *      reference runs one WASM instance per worker for this, SURVEY.md 8b "Threading").  scalars: count * n * scalar_size bytes,
 *      MSM j uses the j-th block; out: count * 3*n8 bytes.  Internally spread over `batch_workers` (option, default 8; each runs its MSMs on one lane, "batch_lanes") sub-contexts
 *      with their own streams, scratch and host threads.  Returns when all results are in `out`. */
int b200msm_g1_multiexp_batch(b200msm_ctx* ctx, uint64_t handle, const void* scalars, uint32_t scalar_size, uint64_t n,
                              uint32_t count, void* out);

/* ---- == g1m_normalize + f1m_fromMontgomery(x), f1m_fromMontgomery(y)   src/build_curve_jacobian_a0.js:940-973,
 *      the comparison form of the reference's tests (test/batchAffine.js:1249-1254):
 * count Jacobian Montgomery points -> count canonical affine points x || y as plain LE integers < q; infinity -> zeros. */
int b200msm_g1_normalize(b200msm_ctx* ctx, int curve, const void* jac, uint64_t count, void* affine_canonical);

/* ---- == repeated g1m_add (src/build_curve_jacobian_a0.js:541-658): out = sum of count Jacobian points.
 * Used to merge the per-GPU partial results of a point-range-sharded MSM. */
int b200msm_g1_sum(b200msm_ctx* ctx, int curve, const void* jac_points, uint64_t count, void* out);

/* ---- point codecs and batch conversions on either side of the MSM (the formats an ffjavascript / snarkjs host holds its
 * points in): op 0 == g1m_batchLEMtoU, 1 == g1m_batchLEMtoC, 2 == g1m_batchUtoLEM, 3 == g1m_batchCtoLEM
 * (src/build_curve_jacobian_a0.js:1166-1328,1413-1416), 4 == g1m_batchToAffine (:1040-1125), 5 == g1m_batchToJacobian (:1418).
 * Element sizes in -> out (bytes): 0: 2n8 -> 2n8, 1: 2n8 -> n8, 2: 2n8 -> 2n8, 3: n8 -> 2n8, 4: 3n8 -> 2n8, 5: 2n8 -> 3n8.
 * "LEM" = little-endian Montgomery affine (the MSM's input format); "U"/"C" = big-endian plain uncompressed/compressed,
 * first byte 0x40 = infinity, 0x80 (compressed only) = y is the greater of the two roots.  The G2 curve ids give g2m_batch* (elements Fq2 = c0 || c1,
 * n8 = 96 / 64; the byte reversal covers the whole element, the sign is f2m_sign, the root f2m_sqrt). */
int b200msm_g1_batch_convert(b200msm_ctx* ctx, int curve, int op, const void* in, uint64_t n, void* out);

/* ---- GLV pre-pass, BLS12-381 only like the reference (src/build_glv.js:3; other curves: B200MSM_E_UNSUPPORTED).
 * == g1m_glv_decomposeScalar(pScalar, pScalarRes) -> sign over a batch                     src/build_glv.js:53-146
 * scalars: n x 32 bytes.  out_scalars: n x 64 bytes = |k1| (bytes 0..15) and |k2| (bytes 32..47), other bytes zero,
 * exactly the reference's pScalarRes; out_signs (nullable): n x uint32, bit 0 = (k1 >= 0), bit 1 = (k2 >= 0). */
int b200msm_glv_decompose_scalars(b200msm_ctx* ctx, int curve, const void* scalars, uint64_t n, void* out_scalars, void* out_signs);
/* == g1m_glv_preprocessEndomorphism(pPoints, pScalars, numPoints, pPointsRes, pScalarsRes)  src/build_glv.js:178-263
 * n affine points (96 B) / n scalars (32 B) -> 2n points [P_i or -P_i, phi(P_i) or -phi(P_i)] and 2n scalars of 32 bytes
 * (each < 2^128); feeding them to b200msm_g1_multiexp_affine(scalar_size = 32, 2n) gives the same sum (test/glv.js:103-192). */
int b200msm_g1_glv_preprocess(b200msm_ctx* ctx, int curve, const void* points, const void* scalars, uint64_t n,
                              void* out_points, void* out_scalars);

/* ---- == frm_fft(px, n) / frm_ifft(px, n) with n = 2^log2n                                    src/build_fft.js:178-245
 * (frm wired at src/bls12381/build_bls12381.js:39-43, src/bn128/build_bn128.js:35-39).  Fr elements: 32 bytes little-endian,
 * Montgomery form a * 2^256 mod r, reduced.  out[k] = sum_j in[j] * w^(j*k) with w = ROOTs[log2n] (build_fft.js:44-63); the
 * inverse is the same transform followed by i -> n - i and the factor 1/n (:396-516).  curve: 0 = BLS12-381 Fr (log2n <= 28),
 * 1 = BN254 Fr (log2n <= 28).  in / out: host or device, n * 32 bytes each; in == out transforms in place. */
int b200msm_fr_fft(b200msm_ctx* ctx, int curve, const void* in, uint32_t log2n, int inverse, void* out);
/* measurement hook: CUDA-event times (ms) of the last transform on this context: bit reversal, shared-memory tile stages, global
 * radix-4/2 passes, inverse finalisation; passes[0] / passes[1] = number of radix-4 / radix-2 global passes it ran. */
int b200msm_fr_fft_last_phases(b200msm_ctx* ctx, float ms[4], uint32_t passes[2]);

/* ---- synthetic inputs (benchmarks/multiexp.js:16-23 builds bases on the module itself):
 * device_out[i] = k_i * G for i in [0, n), affine Montgomery, k_i = splitmix64(seed + first + i) (0 mapped to 1).
 * device_out must be a device pointer with room for n * 2*n8 bytes. */
int b200msm_g1_generate_bases(b200msm_ctx* ctx, int curve, uint64_t seed, uint64_t first, uint64_t n, void* device_out);

/* ---- kernel-level parity hooks: r[i] = op(a[i], b[i]) on Montgomery Fq elements
 * op: 0 f1m_mul  1 f1m_add  2 f1m_sub  3 f1m_square  4 f1m_inverse (the engine's: optimised binary GCD)  5 f1m_toMontgomery
 *     6 f1m_fromMontgomery  7 f1m_neg  8 f1m_inverse by Fermat (cross-check)       (src/build_f1m.js:71-105, 466-777, 779-1076, 1089-1122) */
int b200msm_fq_op(b200msm_ctx* ctx, int curve, int op, const void* a, const void* b, void* r, uint64_t count);

/* ---- f1m_batchInverse (src/build_batchinverse.js:4-140; f2m_batchInverse for the G2 ids): out[i] = 1 / in[i] on Montgomery elements of the
 * curve's coordinate field, zeros stay zero like in the reference.  One field inversion for the whole array: the grid-wide product tree the
 * batch-affine rounds of the MSM use (k_prod_fwd / k_inv_root / k_prod_bwd).  Host or device pointers; in == out is allowed. */
int b200msm_fq_batch_inverse(b200msm_ctx* ctx, int curve, const void* in, uint64_t count, void* out);

/* ---- test hook: the digit / sort phase alone -- computeSchedule + organizeBuckets (src/build_multiexp_opt.js:175-347, 364-633) as the engine
 * runs them (k_digits<count>, scan, k_digits<scatter>).  plan_out = {Wd windows, W bucket slots, B buckets per slot, c0, rem, nbits}: windows
 * 0 .. rem-1 are c0+1 bits wide, the others c0 bits; windows 0 .. Wd-2 use signed digits in [-B, B] (|digit| - 1 = bucket, sign in bit 31 of the
 * entry), the last window is unsigned (digits above B live in slot Wd when there is one).  offsets_out (HOST, W*B + 1 words, slot-major) and
 * sorted_out (HOST, offsets[W*B] words: point index | sign << 31, bucket by bucket; the order inside a bucket is unspecified).
 * window_bits = 0: the engine's own choice for n.  offsets_out == NULL: only the plan is written. */
int b200msm_debug_schedule(b200msm_ctx* ctx, const void* scalars, uint32_t scalar_size, uint64_t n, uint32_t window_bits, uint32_t plan_out[6],
                           uint32_t* offsets_out, uint64_t offsets_cap, uint32_t* sorted_out, uint64_t sorted_cap);

/* ---- tuning knobs (never change results).  key: "window_bits" (0 = auto), "accumulate" (0 = auto, 1 = serial, 2 = batch-affine),
 * "tree_rounds" (-1 = auto), "combine" (0 = serial tail on the host (default), 1 = device chain k_window_sums + k_horner),
 * "lanes" (1..8 overlapping accumulate streams), "issue_threads" (1 = one issuing host thread per lane, 0 = the calling
 * thread issues every lane (default; measured equal, profiles/README.md r2)), "sort_groups" (1 = the sort is pipelined per window group on the lanes' streams (default)), "batch_workers" (8), "batch_lanes" (1), "batch_blocking" (1 = the batch workers sleep in their host waits instead of spinning: set it before the first batch when the
 * workers of several processes outnumber the host cores), "fold_cluster", "multi_min_points", "multi_replicate" (multi-device contexts, see
 * b200msm_create_multi), "xonly" (1 = from 2^19 points on, round 0's forward pass gathers from a copy of the bases' x coordinates alone (default), 0 = from the
 * bases themselves), "group_plan" / "meta_upfront" / "ba_k" / "pt_k" / "persist" / "groups" / "group_small" / "subslots" / "group_pairs" (scheduling shapes measured in
 * profiles/README.md; the defaults are the measured optima).
 * On a multi-device context an option applies to every device. */
int b200msm_set_option(b200msm_ctx* ctx, const char* key, int64_t value);

/* ---- counters since context creation.  key: "launches" (kernel launches issued by this context, all its devices and batch workers). */
int b200msm_get_counter(b200msm_ctx* ctx, const char* key, uint64_t* value);

/* ---- host-only: field constants as the engine uses them (q, R mod q, R^2 mod q as n8-byte LE; np32). No GPU needed. */
int b200msm_constants(int curve, uint32_t* n8, uint8_t* q, uint8_t* r_mod_q, uint8_t* r2_mod_q, uint32_t* np32);

#ifdef __cplusplus
}
#endif
#endif /* B200MSM_H */
