/*
 * oracle.h -- CPU restatement of the curve25519-dalek `serial` u64 backend hot path.
 *
 * TEST INFRASTRUCTURE.  This library is the parity checker for the CUDA engine
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * leg).  It is never linked into, or called from, the product library
 * (curve25519_dalek_b200/csrc).  Every function cites the reference file:line
 * it restates (paths relative to /root/reference/).
 *
 * Abbreviations: C/ = curve25519-dalek/src/, E/ = ed25519-dalek/src/.
 *
 * Parity pinning: field / scalar / Edwards / Ristretto / Ed25519 arithmetic is
 * pinned against the reference's own known-answer vectors (tests/golden/).
 * The Merlin/STROBE-128 transcript (E/batch/transcript.rs, third-party
 * strobe-rs 0.13.0 + keccak 0.2.0, absent from /root/reference) has no golden
 * output in the reference tree: its z_i VALUES are "parity unpinned" against
 * the reference itself (they are pinned against the public merlin
 * conformance vector and hashlib's Keccak); the verify_batch VERDICT is pinned.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- field: C/backend/serial/u64/field.rs ---------------- */
typedef struct { uint64_t v[5]; } fe51;            /* field.rs:43 */

void fe_zero(fe51 *o);
void fe_one(fe51 *o);
void fe_add(fe51 *o, const fe51 *a, const fe51 *b);            /* field.rs:58-73 */
void fe_sub(fe51 *o, const fe51 *a, const fe51 *b);            /* field.rs:82-102 */
void fe_mul(fe51 *o, const fe51 *a, const fe51 *b);            /* field.rs:111-214 */
void fe_neg(fe51 *o, const fe51 *a);                           /* field.rs:276-287 */
void fe_pow2k(fe51 *o, const fe51 *a, uint32_t k);             /* field.rs:454-559 */
void fe_square(fe51 *o, const fe51 *a);                        /* field.rs:562-564 */
void fe_square2(fe51 *o, const fe51 *a);                       /* field.rs:567-574 */
void fe_from_bytes(fe51 *o, const uint8_t s[32]);              /* field.rs:338-363 */
void fe_to_bytes(uint8_t s[32], const fe51 *a);                /* field.rs:368-450 */
int  fe_ct_eq(const fe51 *a, const fe51 *b);                   /* C/field.rs:92-99 */
int  fe_is_negative(const fe51 *a);                            /* C/field.rs:156-159 */
int  fe_is_zero(const fe51 *a);                                /* C/field.rs:166-171 */
void fe_cond_assign(fe51 *o, const fe51 *a, int c);
void fe_cond_negate(fe51 *o, int c);
void fe_invert(fe51 *o, const fe51 *a);                        /* C/field.rs:283-292 */
void fe_pow_p58(fe51 *o, const fe51 *a);                       /* C/field.rs:297-306 */
int  fe_sqrt_ratio_i(fe51 *r, const fe51 *u, const fe51 *v);   /* C/field.rs:320-366 */
int  fe_invsqrt(fe51 *r, const fe51 *a);                       /* C/field.rs:380-382 */
void fe_invert_batch(fe51 *inputs, size_t n);                  /* C/field.rs:239-273 */

/* ---------------- scalars: C/backend/serial/u64/scalar.rs, C/scalar.rs -------- */
typedef struct { uint64_t v[5]; } sc52;            /* u64/scalar.rs:26 */

void sc52_from_bytes(sc52 *o, const uint8_t s[32]);            /* u64/scalar.rs:66-86 */
void sc52_from_bytes_wide(sc52 *o, const uint8_t s[64]);       /* u64/scalar.rs:89-116 */
void sc52_to_bytes(uint8_t s[32], const sc52 *a);              /* u64/scalar.rs:121-158 */
void sc52_add(sc52 *o, const sc52 *a, const sc52 *b);          /* u64/scalar.rs:161-174 */
void sc52_sub(sc52 *o, const sc52 *a, const sc52 *b);          /* u64/scalar.rs:177-191 */
void sc52_mul(sc52 *o, const sc52 *a, const sc52 *b);          /* u64/scalar.rs:302-305 */
void sc52_montgomery_mul(sc52 *o, const sc52 *a, const sc52 *b); /* u64/scalar.rs:317-319 */

/* Scalar = 32 little-endian bytes (C/scalar.rs:193-230) */
void scalar_reduce(uint8_t o[32], const uint8_t a[32]);             /* C/scalar.rs:1153-1160 (reduce), :235-244 */
void scalar_from_bytes_mod_order_wide(uint8_t o[32], const uint8_t a[64]); /* C/scalar.rs:248-250 */
int  scalar_is_canonical(const uint8_t a[32]);                      /* C/scalar.rs:259-263, :1163-1166 */
void scalar_add(uint8_t o[32], const uint8_t a[32], const uint8_t b[32]);  /* C/scalar.rs:340-349 */
void scalar_sub(uint8_t o[32], const uint8_t a[32], const uint8_t b[32]);  /* C/scalar.rs:353-362 */
void scalar_mul(uint8_t o[32], const uint8_t a[32], const uint8_t b[32]);  /* C/scalar.rs:317-322 */
void scalar_neg(uint8_t o[32], const uint8_t a[32]);                /* C/scalar.rs:366-374 */
void scalar_invert(uint8_t o[32], const uint8_t a[32]);             /* C/scalar.rs:739-741 (value only) */
void scalar_invert_batch(uint8_t *inout, size_t n, uint8_t ret[32]);  /* C/scalar.rs:793-853 (values) */
void scalar_from_u64(uint8_t o[32], uint64_t x);
void scalar_non_adjacent_form(int8_t naf[256], const uint8_t a[32], unsigned w); /* C/scalar.rs:955-1007 */
void scalar_as_radix_16(int8_t out[64], const uint8_t a[32]);       /* C/scalar.rs:1019-1051 */
size_t scalar_to_radix_2w_size_hint(unsigned w);                    /* C/scalar.rs:1056-1069 */
void scalar_as_radix_2w(int8_t out[64], const uint8_t a[32], unsigned w); /* C/scalar.rs:1093-1150 */

/* ---------------- curve models: C/backend/serial/curve_models.rs, C/edwards.rs --- */
typedef struct { fe51 X, Y, Z, T; } ge_p3;         /* EdwardsPoint, C/edwards.rs:390-395 */
typedef struct { fe51 X, Y, Z; } ge_p2;            /* ProjectivePoint, curve_models.rs:154 */
typedef struct { fe51 X, Y, Z, T; } ge_p1p1;       /* CompletedPoint, curve_models.rs:169 */
typedef struct { fe51 y_plus_x, y_minus_x, xy2d; } ge_aniels;     /* curve_models.rs:184 */
typedef struct { fe51 Y_plus_X, Y_minus_X, Z, T2d; } ge_pniels;   /* curve_models.rs:206 */

void ge_identity(ge_p3 *o);                                           /* C/edwards.rs:428-437 */
void ge_basepoint(ge_p3 *o);                                          /* u64/constants.rs:163-186 */
void ge_p2_identity(ge_p2 *o);
void ge_p2_to_p3(ge_p3 *o, const ge_p2 *p);                           /* curve_models.rs:338-345 */
void ge_p1p1_to_p2(ge_p2 *o, const ge_p1p1 *p);                       /* curve_models.rs:353-359 */
void ge_p1p1_to_p3(ge_p3 *o, const ge_p1p1 *p);                       /* curve_models.rs:365-372 */
void ge_p2_double(ge_p1p1 *o, const ge_p2 *p);                        /* curve_models.rs:381-397 */
void ge_add_pniels(ge_p1p1 *o, const ge_p3 *p, const ge_pniels *q);   /* curve_models.rs:411-430 */
void ge_sub_pniels(ge_p1p1 *o, const ge_p3 *p, const ge_pniels *q);   /* curve_models.rs:433-452 */
void ge_add_aniels(ge_p1p1 *o, const ge_p3 *p, const ge_aniels *q);   /* curve_models.rs:455-473 */
void ge_sub_aniels(ge_p1p1 *o, const ge_p3 *p, const ge_aniels *q);   /* curve_models.rs:476-494 */
void ge_pniels_neg(ge_pniels *o, const ge_pniels *p);                 /* curve_models.rs:500-511 */
void ge_p3_to_pniels(ge_pniels *o, const ge_p3 *p);                   /* C/edwards.rs:528-535 */
void ge_p3_to_p2(ge_p2 *o, const ge_p3 *p);                           /* C/edwards.rs:541-547 */
void ge_p3_double(ge_p3 *o, const ge_p3 *p);                          /* C/edwards.rs:786-788 */
void ge_p3_add(ge_p3 *o, const ge_p3 *p, const ge_p3 *q);             /* C/edwards.rs:795-800 */
void ge_p3_sub(ge_p3 *o, const ge_p3 *p, const ge_p3 *q);             /* C/edwards.rs:818-823 */
void ge_p3_neg(ge_p3 *o, const ge_p3 *p);                             /* C/edwards.rs:860-871 */
void ge_mul_by_pow_2(ge_p3 *o, const ge_p3 *p, uint32_t k);           /* C/edwards.rs:1370-1380 */
int  ge_p3_ct_eq(const ge_p3 *a, const ge_p3 *b);                     /* C/edwards.rs:501-512 */
int  ge_is_identity(const ge_p3 *a);                                  /* C/traits.rs:41-48 */
int  ge_decompress(ge_p3 *o, const uint8_t s[32]);  /* 1 = Some; C/edwards.rs:211-257 */
void ge_compress(uint8_t s[32], const ge_p3 *p);    /* C/edwards.rs:564-617, edwards/affine.rs:71-75 */
void ge_p3_from_limbs(ge_p3 *o, const uint64_t limbs[20]);
void ge_p3_to_limbs(uint64_t limbs[20], const ge_p3 *p);
/* constant-time variable-base scalar multiplication, C/backend/serial/scalar_mul/variable_base.rs:11-48 */
void ge_scalarmul(ge_p3 *o, const uint8_t scalar[32], const ge_p3 *p);
int  ge_is_small_order(const ge_p3 *p);                               /* C/edwards.rs:1405-1407 */
int  ge_is_torsion_free(const ge_p3 *p);                              /* C/edwards.rs:1435-1437 */

/* window tables, C/window.rs */
typedef struct { ge_pniels t[8]; } ge_lookup_table;                   /* window.rs:47, [P..8P] */
typedef struct { ge_pniels t[8]; } ge_naf_table5;                     /* window.rs:183, [A,3A..15A] */
void ge_lookup_table_from(ge_lookup_table *t, const ge_p3 *p);        /* window.rs:97-105 */
void ge_lookup_table_select(ge_pniels *o, const ge_lookup_table *t, int8_t x); /* window.rs:54-76 */
void ge_naf_table5_from(ge_naf_table5 *t, const ge_p3 *a);            /* window.rs:201-211 */

/* ---------------- multiscalar algorithms --------------------------------- */
/* points are ge_p3; `present[i]==0` models a `None` entry (result: return 0 = None). */
int  msm_pippenger(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points,
                   const uint8_t *present, size_t n);   /* scalar_mul/pippenger.rs:67-160 */
int  msm_straus_vartime(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points,
                        const uint8_t *present, size_t n); /* scalar_mul/straus.rs:159-200 */
void msm_straus_ct(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points, size_t n); /* straus.rs:103-144 */
/* EdwardsPoint trait impls with the reference's size dispatch, C/edwards.rs:966-1031 */
int  edwards_optional_multiscalar_mul(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points,
                                      const uint8_t *present, size_t n);
void edwards_multiscalar_mul(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points, size_t n);
/* aA + bB, scalar_mul/vartime_double_base.rs:23-72 (value) */
void edwards_vartime_double_scalar_mul_basepoint(ge_p3 *o, const uint8_t a[32], const ge_p3 *A,
                                                 const uint8_t b[32]);

/* helpers for tests / bench input synthesis (not reference functions) */
/* scalar_mul/precomputed_straus.rs:57-126; 1 = Some, 0 = None, -1 = more static scalars than points */
int  msm_precomputed_straus(ge_p3 *o, const uint8_t *static_scalars, size_t n_static, const ge_p3 *static_points,
                            size_t n_static_points, const uint8_t *dynamic_scalars, const ge_p3 *dynamic_points,
                            const uint8_t *present, size_t n_dynamic);
void ge_compress_batch(uint8_t *out, const ge_p3 *points, size_t n);            /* C/edwards.rs:633-647 */
void ristretto_double_and_compress_batch(uint8_t *out, const ge_p3 *points, size_t n);  /* C/ristretto.rs:564-646 */
void oracle_points_progression(ge_p3 *out, size_t n, const uint8_t t0[32], const uint8_t q[32]);
int  oracle_msm_compressed(uint8_t out[32], const uint8_t *scalars, const ge_p3 *points, size_t n);
int  oracle_msm_limbs(uint64_t out_limbs[20], const uint8_t *scalars, const ge_p3 *points, size_t n);
void oracle_sum_points(uint8_t out[32], uint64_t out_limbs[20], const ge_p3 *points, size_t n);

/* ---------------- ristretto: C/ristretto.rs -------------------------------- */
int  ristretto_decompress(ge_p3 *o, const uint8_t s[32]);             /* ristretto.rs:266-345 */
void ristretto_compress(uint8_t s[32], const ge_p3 *p);               /* ristretto.rs:500-533 */
int  ristretto_ct_eq(const ge_p3 *a, const ge_p3 *b);                 /* ristretto.rs:815-830 */
/* batch of independent constant-time double-base MSMs a_i*G + b_i*H (config 5):
 * RistrettoPoint::multiscalar_mul([a_i,b_i],[G,H]) -> compress, ristretto.rs:964-977 */
int  ristretto_double_base_batch(uint8_t *out /*n*32*/, const uint8_t *a, const uint8_t *b,
                                 const uint8_t G[32], const uint8_t H[32], size_t n);

/* ---------------- hashing ---------------------------------------------------- */
void sha512(uint8_t out[64], const uint8_t *msg, size_t len);         /* FIPS 180-4 (sha2 0.11.0) */
typedef struct { uint64_t h[8]; uint8_t buf[128]; size_t buflen; uint64_t total; } sha512_ctx;
void sha512_init(sha512_ctx *c);
void sha512_update(sha512_ctx *c, const uint8_t *msg, size_t len);
void sha512_final(sha512_ctx *c, uint8_t out[64]);
void keccak_f1600(uint64_t st[25]);                                   /* keccak 0.2.0 f1600 */

/* STROBE-128 subset used by Merlin (strobe-rs 0.13.0, spec v1.0.2) */
typedef struct { uint8_t st[200]; uint8_t pos, pos_begin, cur_flags; } strobe128;
void strobe128_new(strobe128 *s, const uint8_t *proto, size_t len);
void strobe128_meta_ad(strobe128 *s, const uint8_t *d, size_t len, int more);
void strobe128_ad(strobe128 *s, const uint8_t *d, size_t len, int more);
void strobe128_prf(strobe128 *s, uint8_t *d, size_t len, int more);
void strobe128_key(strobe128 *s, const uint8_t *d, size_t len, int more);

/* Merlin transcript wrapper, E/batch/transcript.rs:39-207 */
typedef struct { strobe128 s; } merlin_transcript;
void merlin_new(merlin_transcript *t, const uint8_t *label, size_t len);               /* :54-61 */
void merlin_append_message(merlin_transcript *t, const uint8_t *label, size_t llen,
                           const uint8_t *msg, size_t mlen);                           /* :69-74 */
void merlin_challenge_bytes(merlin_transcript *t, const uint8_t *label, size_t llen,
                            uint8_t *dest, size_t dlen);                               /* :83-88 */
void merlin_rng_finalize_zero(merlin_transcript *rng, const merlin_transcript *t);     /* :96-100, :157-173 with ZeroRng (E/batch.rs:49-76) */
void merlin_rng_fill(merlin_transcript *rng, uint8_t *dest, size_t dlen);              /* :200-206 */

/* ---------------- ed25519: E/batch.rs, E/verifying.rs, E/signature.rs ----------- */
enum {
    ED_OK = 0, ED_ERR_VERIFY = 1, ED_ERR_ARRAY_LENGTH = 2, ED_ERR_SCALAR_FORMAT = 3,
    ED_ERR_POINT_DECOMPRESSION = 4
};
/* verify_batch, E/batch.rs:146-251.  Keys are given as 32-byte encodings and decompressed
 * first (VerifyingKey::from_bytes, E/verifying.rs:167-175 -> ED_ERR_POINT_DECOMPRESSION).
 * If zs_out != NULL it receives the n 16-byte z_i values drawn from the transcript RNG. */
int ed25519_verify_batch(const uint8_t *const *msgs, const size_t *msg_lens,
                         const uint8_t *sigs /*n*64*/, const uint8_t *pubkeys /*n*32*/,
                         size_t n, uint8_t *zs_out);
/* same, but the batch is cut into consecutive chunks of `chunk` signatures, each with its own
 * faithful transcript, and ONE combined equation is checked (the engine's large-n mode). */
int ed25519_verify_batch_chunked(const uint8_t *const *msgs, const size_t *msg_lens,
                                 const uint8_t *sigs, const uint8_t *pubkeys,
                                 size_t n, size_t chunk, uint8_t *zs_out);
/* single verification, E/verifying.rs:203-219, :496-557 */
int ed25519_verify(const uint8_t *msg, size_t len, const uint8_t sig[64], const uint8_t pk[32]);
/* verify_strict, E/verifying.rs:359-382 */
int ed25519_verify_strict(const uint8_t *msg, size_t len, const uint8_t sig[64], const uint8_t pk[32]);
/* RFC 8032 keygen/sign (E/signing.rs, hazmat.rs) -- only to synthesise test/bench inputs */
void ed25519_public_key(uint8_t pk[32], const uint8_t seed[32]);
void ed25519_sign(uint8_t sig[64], const uint8_t *msg, size_t len, const uint8_t seed[32]);

/* ---------------- parallel.c: persistent worker threads over independent slices (bench infrastructure) ---------- */
typedef struct oracle_pool oracle_pool;
oracle_pool *oracle_pool_create(int threads);
void oracle_pool_destroy(oracle_pool *pool);
int oracle_pool_threads(const oracle_pool *pool);
/* each returns the wall time of the job in seconds */
double oracle_pool_msm(oracle_pool *pool, uint8_t out[32], const uint8_t *scalars, const ge_p3 *points, size_t n, size_t slice);
double oracle_pool_verify_batches(oracle_pool *pool, const uint8_t *msgs, size_t msg_len, const uint8_t *sigs,
                                  const uint8_t *keys, size_t n, size_t batch, int *verdicts);
double oracle_pool_verify_each(oracle_pool *pool, const uint8_t *msgs, size_t msg_len, const uint8_t *sigs,
                               const uint8_t *keys, size_t n, int strict, uint8_t *results);
double oracle_pool_double_base(oracle_pool *pool, uint8_t *out, const uint8_t *a, const uint8_t *b, const uint8_t G[32],
                               const uint8_t H[32], size_t n, int *ok);

#ifdef __cplusplus
}
#endif
#endif
