/* msm_oracle.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C CPU restatement of the reference's G1 MSM path (upstream wasmcurves
 * algorithm that the shipped build/{bls12381,bn128}.wasm contain).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Citations are relative to /root/reference/wasmcurves/src/.
 *
 *   field   : Montgomery Fq, R = 2^(64*n64), fully reduced      build_f1m.js:71-105,466-777
 *   group   : Jacobian a=0: add-2007-bl, madd-2007-bl, dbl-2009-l build_curve_jacobian_a0.js:291-359,541-761
 *   msm     : getChunk / _chunk / _reduceTable / multiexp         build_multiexp.js:25-461
 *   output  : normalize + fromMontgomery                          build_curve_jacobian_a0.js:940-973
 *
 * Parity pinned in tests/test_oracle.py against (1) the reference's golden vectors
 * (tests/golden/) and (2) oracle/_ref = the reference's own WASM compiled natively.
 *
 * Byte formats are the reference's: Fq = n8 bytes LE Montgomery; affine = x||y (inf = zeros);
 * Jacobian = x||y||z (inf: z == 0, canonical zero = (0, R mod q, 0)); scalars = plain LE integers.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;
#define MAXL 6

typedef struct {
  int n64;              /* 6 (BLS12-381) or 4 (BN254) */
  uint64_t q[MAXL];     /* modulus */
  uint64_t one[MAXL];   /* R mod q */
  uint64_t r2[MAXL];    /* R^2 mod q */
  uint64_t np;          /* -q^-1 mod 2^64 */
} field_t;

static const field_t F_BLS = {6,
  {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
  {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
  {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull},
  0x89f3fffcfffcfffdull};
static const field_t F_BN = {4,
  {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull, 0, 0},
  {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full, 0, 0},
  {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full, 0, 0},
  0x87d20782e4866389ull};

static const field_t* field_of(int curve) { return curve == 0 ? &F_BLS : &F_BN; }

typedef uint64_t fe[MAXL];

static int fe_is_zero(const field_t* f, const uint64_t* a) { uint64_t o = 0; for (int i = 0; i < f->n64; i++) o |= a[i]; return o == 0; }
static int fe_eq(const field_t* f, const uint64_t* a, const uint64_t* b) { uint64_t o = 0; for (int i = 0; i < f->n64; i++) o |= a[i] ^ b[i]; return o == 0; }
static void fe_copy(const field_t* f, uint64_t* r, const uint64_t* a) { for (int i = 0; i < f->n64; i++) r[i] = a[i]; }
static void fe_zero(const field_t* f, uint64_t* r) { for (int i = 0; i < f->n64; i++) r[i] = 0; }
static int ge_q(const field_t* f, const uint64_t* a) {           /* int_gte, build_int.js:148-184 */
  for (int i = f->n64 - 1; i >= 0; i--) { if (a[i] > f->q[i]) return 1; if (a[i] < f->q[i]) return 0; }
  return 1;
}
static uint64_t sub_n(int n, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t br = 0; for (int i = 0; i < n; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } return br;
}
static uint64_t add_n(int n, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t c = 0; for (int i = 0; i < n; i++) { u128 s = (u128)a[i] + b[i] + c; r[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } return c;
}
/* f1m_add, build_f1m.js:71-89: add, subtract q if carry or >= q */
static void fe_add(const field_t* f, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t c = add_n(f->n64, r, a, b);
  if (c || ge_q(f, r)) sub_n(f->n64, r, r, f->q);
}
/* f1m_sub, build_f1m.js:91-105: subtract, add q on borrow */
static void fe_sub(const field_t* f, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  if (sub_n(f->n64, r, a, b)) add_n(f->n64, r, r, f->q);
}
/* f1m_mul, build_f1m.js:466-777: CIOS Montgomery product with final conditional subtract (:754-765).
   Restated on 64-bit words (the reference interleaves 32-bit limbs in i64 registers; same value). */
static void fe_mul(const field_t* f, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  int n = f->n64; uint64_t t[MAXL + 2] = {0};
  for (int i = 0; i < n; i++) {
    u128 c = 0;
    for (int j = 0; j < n; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[n]; t[n] = (uint64_t)c; t[n + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * f->np;
    c = ((u128)m * f->q[0] + t[0]) >> 64;
    for (int j = 1; j < n; j++) { c += (u128)m * f->q[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[n]; t[n - 1] = (uint64_t)c; t[n] = t[n + 1] + (uint64_t)(c >> 64);
  }
  if (t[n] || ge_q(f, t)) sub_n(n, t, t, f->q);
  for (int i = 0; i < n; i++) r[i] = t[i];
}
static void fe_sqr(const field_t* f, uint64_t* r, const uint64_t* a) { fe_mul(f, r, a, a); }   /* f1m_square :779-1076 (same value) */
static void fe_to_mont(const field_t* f, uint64_t* r, const uint64_t* a) { fe_mul(f, r, a, f->r2); }                 /* :1089 */
static void fe_from_mont(const field_t* f, uint64_t* r, const uint64_t* a) { fe one1 = {1}; fe_mul(f, r, a, one1); } /* :1098 */
/* f1m_inverse, build_f1m.js:1112-1122.  The reference uses extended Euclid (build_int.js:922-1064);
   the value a^-1 mod q is unique, we use Fermat a^(q-2). */
static void fe_inv(const field_t* f, uint64_t* r, const uint64_t* a) {
  uint64_t e[MAXL]; fe two = {2}; sub_n(f->n64, e, f->q, two);
  fe acc; fe_copy(f, acc, f->one);
  for (int i = f->n64 * 64 - 1; i >= 0; i--) {
    fe_sqr(f, acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(f, acc, acc, a);
  }
  fe_copy(f, r, acc);
}

/* ---- Jacobian points: X,Y,Z each n64 words, stored contiguously (3*n8 bytes) */
typedef struct { fe x, y, z; } jac;

static void load_fe(const field_t* f, uint64_t* r, const uint8_t* p) { memcpy(r, p, f->n64 * 8); }
static void store_fe(const field_t* f, uint8_t* p, const uint64_t* a) { memcpy(p, a, f->n64 * 8); }

static void jac_zero(const field_t* f, jac* r) { fe_zero(f, r->x); fe_copy(f, r->y, f->one); fe_zero(f, r->z); }  /* g1m_zero :124-150 */
static int jac_is_zero(const field_t* f, const jac* p) { return fe_is_zero(f, p->z); }                             /* g1m_isZero :40-53 */

/* g1m_double, build_curve_jacobian_a0.js:291-359 (dbl-2009-l, a = 0) */
static void jac_double(const field_t* f, jac* r, const jac* p) {
  if (jac_is_zero(f, p)) { *r = *p; return; }
  fe A, B, C, D, E, Fv, G, x3, y3, z3, e8;
  fe_sqr(f, A, p->x); fe_sqr(f, B, p->y); fe_sqr(f, C, B);
  fe_add(f, D, p->x, B); fe_sqr(f, D, D); fe_sub(f, D, D, A); fe_sub(f, D, D, C); fe_add(f, D, D, D);
  fe_add(f, E, A, A); fe_add(f, E, E, A); fe_sqr(f, Fv, E);
  fe_mul(f, G, p->y, p->z);
  fe_add(f, x3, D, D); fe_sub(f, x3, Fv, x3);
  fe_add(f, e8, C, C); fe_add(f, e8, e8, e8); fe_add(f, e8, e8, e8);
  fe_sub(f, y3, D, x3); fe_mul(f, y3, y3, E); fe_sub(f, y3, y3, e8);
  fe_add(f, z3, G, G);
  fe_copy(f, r->x, x3); fe_copy(f, r->y, y3); fe_copy(f, r->z, z3);
}

/* g1m_add, build_curve_jacobian_a0.js:541-658 (add-2007-bl; inf and equal-point dispatch) */
static void jac_add(const field_t* f, jac* r, const jac* p, const jac* q) {
  if (jac_is_zero(f, p)) { *r = *q; return; }
  if (jac_is_zero(f, q)) { *r = *p; return; }
  fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, x3, y3, z3;
  fe_sqr(f, Z1Z1, p->z); fe_sqr(f, Z2Z2, q->z);
  fe_mul(f, U1, p->x, Z2Z2); fe_mul(f, U2, q->x, Z1Z1);
  fe_mul(f, S1, p->y, q->z); fe_mul(f, S1, S1, Z2Z2);
  fe_mul(f, S2, q->y, p->z); fe_mul(f, S2, S2, Z1Z1);
  if (fe_eq(f, U1, U2)) {
    if (fe_eq(f, S1, S2)) { jac_double(f, r, p); return; }
    /* P + (-P): H = 0 => Z3 = 0 below (a representation of infinity); make it canonical */
    jac_zero(f, r); return;
  }
  fe_sub(f, H, U2, U1);
  fe_add(f, I, H, H); fe_sqr(f, I, I);
  fe_mul(f, J, H, I);
  fe_sub(f, rr, S2, S1); fe_add(f, rr, rr, rr);
  fe_mul(f, V, U1, I);
  fe_sqr(f, x3, rr); fe_sub(f, x3, x3, J); fe_sub(f, x3, x3, V); fe_sub(f, x3, x3, V);
  fe_sub(f, y3, V, x3); fe_mul(f, y3, y3, rr); fe_mul(f, t, S1, J); fe_add(f, t, t, t); fe_sub(f, y3, y3, t);
  fe_add(f, z3, p->z, q->z); fe_sqr(f, z3, z3); fe_sub(f, z3, z3, Z1Z1); fe_sub(f, z3, z3, Z2Z2); fe_mul(f, z3, z3, H);
  fe_copy(f, r->x, x3); fe_copy(f, r->y, y3); fe_copy(f, r->z, z3);
}

/* g1m_addMixed, build_curve_jacobian_a0.js:661-761 (madd-2007-bl).  q is affine x||y, inf = (0,0) (:705-711) */
static void jac_add_mixed(const field_t* f, jac* r, const jac* p, const uint64_t* qx, const uint64_t* qy) {
  if (fe_is_zero(f, qx) && fe_is_zero(f, qy)) { *r = *p; return; }
  if (jac_is_zero(f, p)) { fe_copy(f, r->x, qx); fe_copy(f, r->y, qy); fe_copy(f, r->z, f->one); return; }
  fe Z1Z1, U2, S2, H, HH, I, J, rr, V, t, x3, y3, z3;
  fe_sqr(f, Z1Z1, p->z);
  fe_mul(f, U2, qx, Z1Z1);
  fe_mul(f, S2, qy, p->z); fe_mul(f, S2, S2, Z1Z1);
  if (fe_eq(f, p->x, U2)) {
    if (fe_eq(f, p->y, S2)) { jac_double(f, r, p); return; }
    jac_zero(f, r); return;
  }
  fe_sub(f, H, U2, p->x); fe_sqr(f, HH, H);
  fe_add(f, I, HH, HH); fe_add(f, I, I, I);
  fe_mul(f, J, H, I);
  fe_sub(f, rr, S2, p->y); fe_add(f, rr, rr, rr);
  fe_mul(f, V, p->x, I);
  fe_sqr(f, x3, rr); fe_sub(f, x3, x3, J); fe_sub(f, x3, x3, V); fe_sub(f, x3, x3, V);
  fe_sub(f, y3, V, x3); fe_mul(f, y3, y3, rr); fe_mul(f, t, p->y, J); fe_add(f, t, t, t); fe_sub(f, y3, y3, t);
  fe_add(f, z3, p->z, H); fe_sqr(f, z3, z3); fe_sub(f, z3, z3, Z1Z1); fe_sub(f, z3, z3, HH);
  fe_copy(f, r->x, x3); fe_copy(f, r->y, y3); fe_copy(f, r->z, z3);
}

static void jac_store(const field_t* f, uint8_t* out, const jac* p) {
  int n8 = f->n64 * 8; store_fe(f, out, p->x); store_fe(f, out + n8, p->y); store_fe(f, out + 2 * n8, p->z);
}
static void jac_load(const field_t* f, jac* p, const uint8_t* in) {
  int n8 = f->n64 * 8; memset(p, 0, sizeof *p); load_fe(f, p->x, in); load_fe(f, p->y, in + n8); load_fe(f, p->z, in + 2 * n8);
}

/* _getChunk, build_multiexp.js:25-94.  The reference over-reads up to 3 bytes past the scalar
   (SURVEY 8a defect 4); the mask makes those bits irrelevant, so we read them as zero. */
static uint32_t get_chunk(const uint8_t* s, uint32_t scalar_size, uint32_t start_bit, uint32_t chunk) {
  int32_t bits_to_end = (int32_t)(scalar_size * 8 - start_bit);
  uint32_t mask = ((int32_t)chunk > bits_to_end) ? ((1u << bits_to_end) - 1) : ((1u << chunk) - 1);
  uint32_t o = start_bit >> 3, w = 0;
  for (uint32_t k = 0; k < 4 && o + k < scalar_size; k++) w |= (uint32_t)s[o + k] << (8 * k);
  return (w >> (start_bit & 7)) & mask;
}

/* _reduceTable, build_multiexp.js:373-461 */
static void reduce_table(const field_t* f, jac* t, int p) {
  if (p == 1) return;
  uint32_t half = 1u << (p - 1);
  jac* acc = &t[half - 1];
  for (uint32_t i = 0; i + 1 < half; i++) { jac_add(f, &t[i], &t[i], &t[half + i]); jac_add(f, acc, acc, &t[half + i]); }
  reduce_table(f, t, p - 1);
  for (int k = p - 1; k > 0; k--) jac_double(f, acc, acc);
  jac_add(f, &t[0], &t[0], acc);
}

/* g1m_multiexpAffine_chunk, build_multiexp.js:96-249 */
static void chunk_impl(const field_t* f, const uint8_t* bases, const uint8_t* scalars, uint32_t ssz, uint64_t n,
                       uint32_t start_bit, uint32_t c, jac* out) {
  if (n == 0) { jac_zero(f, out); return; }
  int n8 = f->n64 * 8; uint32_t nt = 1u << c;
  jac* t = (jac*)malloc(sizeof(jac) * nt);
  for (uint32_t j = 0; j < nt; j++) jac_zero(f, &t[j]);
  for (uint64_t i = 0; i < n; i++) {
    uint32_t idx = get_chunk(scalars + i * ssz, ssz, start_bit, c);
    if (idx) { fe x = {0}, y = {0}; load_fe(f, x, bases + i * 2 * n8); load_fe(f, y, bases + i * 2 * n8 + n8); jac_add_mixed(f, &t[idx - 1], &t[idx - 1], x, y); }
  }
  reduce_table(f, t, (int)c);
  *out = t[0];
  free(t);
}

static const uint8_t TSIZES[32] = {17, 17, 17, 17, 17, 17, 17, 17, 17, 17, 16, 16, 15, 14, 13, 13,
                                   12, 11, 10, 9, 8, 7, 7, 6, 5, 4, 3, 2, 1, 1, 1, 1};  /* build_multiexp.js:275-280 */

/* ------------------------------------------------------------------ exported API (ctypes) */

int oracle_multiexp_affine_chunk(int curve, const uint8_t* bases, const uint8_t* scalars, uint32_t ssz, uint64_t n,
                                 uint32_t start_bit, uint32_t c, uint8_t* out_jac) {
  const field_t* f = field_of(curve); jac r; chunk_impl(f, bases, scalars, ssz, n, start_bit, c, &r); jac_store(f, out_jac, &r); return 0;
}

/* g1m_multiexpAffine, build_multiexp.js:251-371 */
int oracle_multiexp_affine(int curve, const uint8_t* bases, const uint8_t* scalars, uint32_t ssz, uint64_t n, uint8_t* out_jac) {
  const field_t* f = field_of(curve); jac pr, aux; jac_zero(f, &pr);
  if (n) {
    uint32_t c = TSIZES[__builtin_clz((uint32_t)n)];
    uint32_t nchunks = (ssz * 8 - 1) / c + 1;
    for (int32_t bit = (int32_t)((nchunks - 1) * c); bit >= 0; bit -= (int32_t)c) {
      if (!jac_is_zero(f, &pr)) for (uint32_t j = 0; j < c; j++) jac_double(f, &pr, &pr);
      chunk_impl(f, bases, scalars, ssz, n, (uint32_t)bit, c, &aux);
      jac_add(f, &pr, &pr, &aux);
    }
  }
  jac_store(f, out_jac, &pr); return 0;
}

/* g1m_normalize + f1m_fromMontgomery x2 -> canonical affine x||y (plain LE integers), inf = zeros.
   build_curve_jacobian_a0.js:940-973; test/batchAffine.js:1249-1254 */
int oracle_normalize(int curve, const uint8_t* in_jac, uint8_t* out_xy) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8; jac p; jac_load(f, &p, in_jac);
  memset(out_xy, 0, 2 * n8);
  if (jac_is_zero(f, &p)) return 0;
  fe zi, zi2, zi3, x, y;
  fe_inv(f, zi, p.z); fe_sqr(f, zi2, zi); fe_mul(f, zi3, zi2, zi);
  fe_mul(f, x, p.x, zi2); fe_mul(f, y, p.y, zi3);
  fe_from_mont(f, x, x); fe_from_mont(f, y, y);
  store_fe(f, out_xy, x); store_fe(f, out_xy + n8, y); return 0;
}

int oracle_add(int curve, const uint8_t* a_jac, const uint8_t* b_jac, uint8_t* out_jac) {
  const field_t* f = field_of(curve); jac a, b, r; jac_load(f, &a, a_jac); jac_load(f, &b, b_jac); jac_add(f, &r, &a, &b); jac_store(f, out_jac, &r); return 0;
}

/* k * (affine Montgomery point), k = LE integer of kbytes bytes: double-and-add (build_timesscalar.js:20-87 semantics) */
int oracle_times_scalar_affine(int curve, const uint8_t* base_xy, const uint8_t* k, uint32_t kbytes, uint8_t* out_jac) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8; jac r; jac_zero(f, &r);
  fe x = {0}, y = {0}; load_fe(f, x, base_xy); load_fe(f, y, base_xy + n8);
  for (int i = (int)kbytes * 8 - 1; i >= 0; i--) {
    jac_double(f, &r, &r);
    if ((k[i >> 3] >> (i & 7)) & 1) jac_add_mixed(f, &r, &r, x, y);
  }
  jac_store(f, out_jac, &r); return 0;
}

/* Synthetic input generator shared by tests and bench: P_i = k_i * G, affine Montgomery,
   k_i = splitmix64(seed + i) (never 0 mod r in practice; guarded).  g is the generator x||y (Montgomery). */
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
uint64_t oracle_point_scalar(uint64_t seed, uint64_t i) { uint64_t k = splitmix64(seed + i); return k ? k : 1; }

int oracle_generate_bases(int curve, const uint8_t* gen_xy, uint64_t seed, uint64_t first, uint64_t n, uint8_t* out_xy) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  jac* pts = (jac*)malloc(sizeof(jac) * (n ? n : 1));
  fe* pre = (fe*)malloc(sizeof(fe) * (n ? n : 1));
  for (uint64_t i = 0; i < n; i++) {
    uint64_t k = oracle_point_scalar(seed, first + i); uint8_t kb[8]; memcpy(kb, &k, 8);
    uint8_t tmp[3 * 48]; oracle_times_scalar_affine(curve, gen_xy, kb, 8, tmp); jac_load(f, &pts[i], tmp);
  }
  /* g1m_batchToAffine, build_curve_jacobian_a0.js:1040-1125: Montgomery-trick batch inversion of z (build_batchinverse.js:4-140) */
  fe acc; fe_copy(f, acc, f->one);
  for (uint64_t i = 0; i < n; i++) { fe_copy(f, pre[i], acc); if (!jac_is_zero(f, &pts[i])) fe_mul(f, acc, acc, pts[i].z); }
  fe inv; fe_inv(f, inv, acc);
  for (uint64_t i = n; i-- > 0;) {
    uint8_t* o = out_xy + i * 2 * n8;
    if (jac_is_zero(f, &pts[i])) { memset(o, 0, 2 * n8); continue; }
    fe zi, zi2, zi3, x, y; fe_mul(f, zi, inv, pre[i]); fe_mul(f, inv, inv, pts[i].z);
    fe_sqr(f, zi2, zi); fe_mul(f, zi3, zi2, zi); fe_mul(f, x, pts[i].x, zi2); fe_mul(f, y, pts[i].y, zi3);
    store_fe(f, o, x); store_fe(f, o + n8, y);
  }
  free(pts); free(pre); return 0;
}

/* field helpers for kernel-level parity tests */
int oracle_fe_mul(int curve, const uint8_t* a, const uint8_t* b, uint8_t* r, uint64_t n) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  for (uint64_t i = 0; i < n; i++) { fe x = {0}, y = {0}, z; load_fe(f, x, a + i * n8); load_fe(f, y, b + i * n8); fe_mul(f, z, x, y); store_fe(f, r + i * n8, z); }
  return 0;
}
int oracle_fe_add(int curve, const uint8_t* a, const uint8_t* b, uint8_t* r, uint64_t n) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  for (uint64_t i = 0; i < n; i++) { fe x = {0}, y = {0}, z; load_fe(f, x, a + i * n8); load_fe(f, y, b + i * n8); fe_add(f, z, x, y); store_fe(f, r + i * n8, z); }
  return 0;
}
int oracle_fe_sub(int curve, const uint8_t* a, const uint8_t* b, uint8_t* r, uint64_t n) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  for (uint64_t i = 0; i < n; i++) { fe x = {0}, y = {0}, z; load_fe(f, x, a + i * n8); load_fe(f, y, b + i * n8); fe_sub(f, z, x, y); store_fe(f, r + i * n8, z); }
  return 0;
}
int oracle_fe_inv(int curve, const uint8_t* a, uint8_t* r, uint64_t n) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  for (uint64_t i = 0; i < n; i++) { fe x = {0}, z; load_fe(f, x, a + i * n8); if (fe_is_zero(f, x)) fe_zero(f, z); else fe_inv(f, z, x); store_fe(f, r + i * n8, z); }
  return 0;
}
int oracle_fe_to_mont(int curve, const uint8_t* a, uint8_t* r, uint64_t n) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  for (uint64_t i = 0; i < n; i++) { fe x = {0}, z; load_fe(f, x, a + i * n8); fe_to_mont(f, z, x); store_fe(f, r + i * n8, z); }
  return 0;
}
int oracle_fe_from_mont(int curve, const uint8_t* a, uint8_t* r, uint64_t n) {
  const field_t* f = field_of(curve); int n8 = f->n64 * 8;
  for (uint64_t i = 0; i < n; i++) { fe x = {0}, z; load_fe(f, x, a + i * n8); fe_from_mont(f, z, x); store_fe(f, r + i * n8, z); }
  return 0;
}
