/*
 * msm.c -- Pippenger and Straus multiscalar multiplication, with the
 * EdwardsPoint trait dispatch.  TEST INFRASTRUCTURE (oracle).
 * Restates C/backend/serial/scalar_mul/{pippenger,straus,vartime_double_base,precomputed_straus}.rs,
 * C/edwards.rs:966-1031 and EdwardsPoint::compress_batch (C/edwards.rs:619-647).
 */
#include "oracle.h"
#include "constants.h"
#include <stdlib.h>
#include <string.h>

static void fe_const(fe51 *o, const uint64_t k[5]) { memcpy(o->v, k, sizeof o->v); }

/* scalar_mul/pippenger.rs:67-160 */
int msm_pippenger(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points,
                  const uint8_t *present, size_t n)
{
    unsigned w = n < 500 ? 6 : (n < 800 ? 7 : 8);                 /* :81-87 */
    size_t max_digit = (size_t)1 << w;
    size_t digits_count = scalar_to_radix_2w_size_hint(w);        /* :90 */
    size_t buckets_count = max_digit / 2;                         /* :91 */

    for (size_t i = 0; i < n; i++) if (present && !present[i]) return 0;   /* :101-104 None */

    int8_t *digits = (int8_t *)malloc(64 * (n ? n : 1));
    ge_pniels *pts = (ge_pniels *)malloc(sizeof(ge_pniels) * (n ? n : 1));
    ge_p3 *buckets = (ge_p3 *)malloc(sizeof(ge_p3) * buckets_count);
    for (size_t i = 0; i < n; i++) {
        scalar_as_radix_2w(digits + 64 * i, scalars + 32 * i, w); /* :95 */
        ge_p3_to_pniels(&pts[i], &points[i]);                     /* :97-99 */
    }

    ge_p3 total; int have_total = 0;
    for (size_t di = digits_count; di-- > 0;) {                   /* :112, high column first */
        for (size_t b = 0; b < buckets_count; b++) ge_identity(&buckets[b]);   /* :114-116 */
        for (size_t i = 0; i < n; i++) {                          /* :122-136 */
            int16_t digit = (int16_t)digits[64 * i + di];
            ge_p1p1 r;
            if (digit > 0) {
                size_t b = (size_t)(digit - 1);
                ge_add_pniels(&r, &buckets[b], &pts[i]); ge_p1p1_to_p3(&buckets[b], &r);
            } else if (digit < 0) {
                size_t b = (size_t)(-digit - 1);
                ge_sub_pniels(&r, &buckets[b], &pts[i]); ge_p1p1_to_p3(&buckets[b], &r);
            }
        }
        ge_p3 isum = buckets[buckets_count - 1], sum = buckets[buckets_count - 1];   /* :146-147 */
        for (size_t i = buckets_count - 1; i-- > 0;) {            /* :148-151 */
            ge_p3_add(&isum, &isum, &buckets[i]);
            ge_p3_add(&sum, &sum, &isum);
        }
        if (!have_total) { total = sum; have_total = 1; }         /* :157 hi_column */
        else { ge_p3 t; ge_mul_by_pow_2(&t, &total, w); ge_p3_add(&total, &t, &sum); }  /* :159 */
    }
    free(digits); free(pts); free(buckets);
    *o = total;
    return 1;
}

/* scalar_mul/straus.rs:159-200 */
int msm_straus_vartime(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points,
                       const uint8_t *present, size_t n)
{
    for (size_t i = 0; i < n; i++) if (present && !present[i]) return 0;   /* :176-179 None */
    int8_t *nafs = (int8_t *)malloc(256 * (n ? n : 1));
    ge_naf_table5 *tabs = (ge_naf_table5 *)malloc(sizeof(ge_naf_table5) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        scalar_non_adjacent_form(nafs + 256 * i, scalars + 32 * i, 5);    /* :171-174 */
        ge_naf_table5_from(&tabs[i], &points[i]);
    }
    ge_p2 r; ge_p2_identity(&r);                                  /* :181 */
    for (int i = 255; i >= 0; i--) {                              /* :183-197 */
        ge_p1p1 t; ge_p3 e;
        ge_p2_double(&t, &r);
        for (size_t j = 0; j < n; j++) {
            int8_t d = nafs[256 * j + i];
            if (d > 0) { ge_p1p1_to_p3(&e, &t); ge_add_pniels(&t, &e, &tabs[j].t[d / 2]); }
            else if (d < 0) { ge_p1p1_to_p3(&e, &t); ge_sub_pniels(&t, &e, &tabs[j].t[(-d) / 2]); }
        }
        ge_p1p1_to_p2(&r, &t);
    }
    ge_p2_to_p3(o, &r);                                           /* :199 */
    free(nafs); free(tabs);
    return 1;
}

/* scalar_mul/straus.rs:103-144 */
void msm_straus_ct(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points, size_t n)
{
    ge_lookup_table *tabs = (ge_lookup_table *)malloc(sizeof(ge_lookup_table) * (n ? n : 1));
    int8_t *digits = (int8_t *)malloc(64 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        ge_lookup_table_from(&tabs[i], &points[i]);               /* :114-117 */
        scalar_as_radix_16(digits + 64 * i, scalars + 32 * i);    /* :123-126 */
    }
    ge_p3 Q; ge_identity(&Q);
    for (int j = 63; j >= 0; j--) {                               /* :129-138 */
        ge_mul_by_pow_2(&Q, &Q, 4);
        for (size_t i = 0; i < n; i++) {
            ge_pniels R; ge_p1p1 t;
            ge_lookup_table_select(&R, &tabs[i], digits[64 * i + j]);
            ge_add_pniels(&t, &Q, &R); ge_p1p1_to_p3(&Q, &t);
        }
    }
    memset(digits, 0, 64 * (n ? n : 1));                          /* :140-141 zeroize */
    free(tabs); free(digits);
    *o = Q;
}

/* C/edwards.rs:1002-1030: size < 190 -> Straus, else Pippenger */
int edwards_optional_multiscalar_mul(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points,
                                     const uint8_t *present, size_t n)
{
    if (n < 190) return msm_straus_vartime(o, scalars, points, present, n);
    return msm_pippenger(o, scalars, points, present, n);
}

/* C/edwards.rs:970-995: always constant-time Straus */
void edwards_multiscalar_mul(ge_p3 *o, const uint8_t *scalars, const ge_p3 *points, size_t n)
{
    msm_straus_ct(o, scalars, points, n);
}

/* value of scalar_mul/vartime_double_base.rs:23-72 (aA + bB); computed with the vartime
 * Straus over [A, B], which yields the same group element */
void edwards_vartime_double_scalar_mul_basepoint(ge_p3 *o, const uint8_t a[32], const ge_p3 *A,
                                                 const uint8_t b[32])
{
    uint8_t sc[64]; ge_p3 pts[2];
    memcpy(sc, a, 32); memcpy(sc + 32, b, 32);
    pts[0] = *A; ge_basepoint(&pts[1]);
    msm_straus_vartime(o, sc, pts, NULL, 2);
}

/* ---- helpers for tests / bench input synthesis (not part of the reference) ---- */
/* out[i] = (t0 + i*q) * B, by one scalar multiplication and repeated addition; the points carry
 * non-trivial Z like the outputs of the reference's own arithmetic. */
void oracle_points_progression(ge_p3 *out, size_t n, const uint8_t t0[32], const uint8_t q[32])
{
    ge_p3 B, Q, P;
    ge_basepoint(&B);
    ge_scalarmul(&P, t0, &B);
    ge_scalarmul(&Q, q, &B);
    for (size_t i = 0; i < n; i++) { out[i] = P; ge_p3_add(&P, &P, &Q); }
}

/* compress(sum scalars[i] * points[i]) with the reference's dispatch; points as n x 20 u64 limbs */
int oracle_msm_compressed(uint8_t out[32], const uint8_t *scalars, const ge_p3 *points, size_t n)
{
    ge_p3 r;
    if (!edwards_optional_multiscalar_mul(&r, scalars, points, NULL, n)) return 0;
    ge_compress(out, &r);
    return 1;
}

/* sum of already-computed points given as limbs (to combine per-thread partial MSMs) */
void oracle_sum_points(uint8_t out[32], uint64_t out_limbs[20], const ge_p3 *points, size_t n)
{
    ge_p3 acc; ge_identity(&acc);
    for (size_t i = 0; i < n; i++) ge_p3_add(&acc, &acc, &points[i]);
    ge_compress(out, &acc);
    if (out_limbs) ge_p3_to_limbs(out_limbs, &acc);
}

int oracle_msm_limbs(uint64_t out_limbs[20], const uint8_t *scalars, const ge_p3 *points, size_t n)
{
    ge_p3 r;
    if (!edwards_optional_multiscalar_mul(&r, scalars, points, NULL, n)) return 0;
    ge_p3_to_limbs(out_limbs, &r);
    return 1;
}

/* ---- VartimePrecomputedMultiscalarMul, serial backend ------------------------------------------------- */
/* C/edwards.rs:551-561 */
static void ge_p3_to_aniels(ge_aniels *o, const ge_p3 *p)
{
    fe51 recip, x, y, d2;
    fe_const(&d2, K_EDWARDS_D2);
    fe_invert(&recip, &p->Z);
    fe_mul(&x, &p->X, &recip);
    fe_mul(&y, &p->Y, &recip);
    fe_add(&o->y_plus_x, &y, &x);
    fe_sub(&o->y_minus_x, &y, &x);
    fe_mul(&o->xy2d, &x, &y); fe_mul(&o->xy2d, &o->xy2d, &d2);
}

/* C/window.rs:266-276: [A, 3A, 5A, ..., 127A] as affine Niels points */
typedef struct { ge_aniels t[64]; } ge_naf_table8;
static void ge_naf_table8_from(ge_naf_table8 *t, const ge_p3 *a)
{
    ge_p3 a2;
    ge_p3_to_aniels(&t->t[0], a);
    ge_p3_double(&a2, a);
    for (int i = 0; i < 63; i++) {
        ge_p1p1 r; ge_p3 e;
        ge_add_aniels(&r, &a2, &t->t[i]); ge_p1p1_to_p3(&e, &r); ge_p3_to_aniels(&t->t[i + 1], &e);
    }
}

/* scalar_mul/precomputed_straus.rs:57-126 (the tables of :37-46 are rebuilt per call: the oracle keeps no state).
 * n_static_points >= n_static (unused points are ignored, :88); a missing dynamic point gives None (:78-81). */
int msm_precomputed_straus(ge_p3 *o, const uint8_t *static_scalars, size_t n_static, const ge_p3 *static_points,
                           size_t n_static_points, const uint8_t *dynamic_scalars, const ge_p3 *dynamic_points,
                           const uint8_t *present, size_t n_dynamic)
{
    if (n_static > n_static_points) return -1;                    /* :88 assert */
    for (size_t i = 0; i < n_dynamic; i++) if (present && !present[i]) return 0;
    int8_t *snaf = (int8_t *)malloc(256 * (n_static ? n_static : 1));
    int8_t *dnaf = (int8_t *)malloc(256 * (n_dynamic ? n_dynamic : 1));
    ge_naf_table8 *stab = (ge_naf_table8 *)malloc(sizeof(ge_naf_table8) * (n_static ? n_static : 1));
    ge_naf_table5 *dtab = (ge_naf_table5 *)malloc(sizeof(ge_naf_table5) * (n_dynamic ? n_dynamic : 1));
    for (size_t i = 0; i < n_static; i++) {
        scalar_non_adjacent_form(snaf + 256 * i, static_scalars + 32 * i, 8);    /* :70-73 */
        ge_naf_table8_from(&stab[i], &static_points[i]);
    }
    for (size_t i = 0; i < n_dynamic; i++) {
        scalar_non_adjacent_form(dnaf + 256 * i, dynamic_scalars + 32 * i, 5);   /* :74-77 */
        ge_naf_table5_from(&dtab[i], &dynamic_points[i]);
    }
    ge_p2 s; ge_p2_identity(&s);                                  /* :94 */
    for (int j = 255; j >= 0; j--) {                              /* :95-123 */
        ge_p1p1 r; ge_p3 e;
        ge_p2_double(&r, &s);
        for (size_t i = 0; i < n_dynamic; i++) {
            int8_t d = dnaf[256 * i + j];
            if (d > 0) { ge_p1p1_to_p3(&e, &r); ge_add_pniels(&r, &e, &dtab[i].t[d / 2]); }
            else if (d < 0) { ge_p1p1_to_p3(&e, &r); ge_sub_pniels(&r, &e, &dtab[i].t[(-d) / 2]); }
        }
        for (size_t i = 0; i < n_static; i++) {
            int8_t d = snaf[256 * i + j];
            if (d > 0) { ge_p1p1_to_p3(&e, &r); ge_add_aniels(&r, &e, &stab[i].t[d / 2]); }
            else if (d < 0) { ge_p1p1_to_p3(&e, &r); ge_sub_aniels(&r, &e, &stab[i].t[(-d) / 2]); }
        }
        ge_p1p1_to_p2(&s, &r);
    }
    ge_p2_to_p3(o, &s);                                           /* :125 */
    free(snaf); free(dnaf); free(stab); free(dtab);
    return 1;
}

/* ---- batch codecs ---------------------------------------------------------------------------------------- */
/* EdwardsPoint::compress_batch_alloc, C/edwards.rs:633-647 */
void ge_compress_batch(uint8_t *out, const ge_p3 *points, size_t n)
{
    fe51 *zs = (fe51 *)malloc(sizeof(fe51) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) zs[i] = points[i].Z;
    fe_invert_batch(zs, n);
    for (size_t i = 0; i < n; i++) {
        fe51 x, y;
        fe_mul(&x, &points[i].X, &zs[i]);
        fe_mul(&y, &points[i].Y, &zs[i]);
        fe_to_bytes(out + 32 * i, &y);                            /* edwards/affine.rs:71-75 */
        out[32 * i + 31] ^= (uint8_t)(fe_is_negative(&x) << 7);
    }
    free(zs);
}
