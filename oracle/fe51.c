/*
 * fe51.c -- GF(2^255-19) in radix 2^51, five u64 limbs, u128 products.
 * TEST INFRASTRUCTURE (oracle).  Restates C/backend/serial/u64/field.rs and the
 * backend-generic helpers of C/field.rs with the same formulas in the same order,
 * so internal limbs equal the reference's, not only the canonical encodings.
 */
#include "oracle.h"
#include "constants.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define MASK51 ((UINT64_C(1) << 51) - 1)

static inline u128 m(uint64_t x, uint64_t y) { return (u128)x * (u128)y; }

void fe_zero(fe51 *o) { memset(o, 0, sizeof *o); }
void fe_one(fe51 *o) { memset(o, 0, sizeof *o); o->v[0] = 1; }

/* field.rs:290-323 `reduce`: weak reduction, carries computed in parallel */
static void fe_reduce(fe51 *o, const uint64_t in[5])
{
    uint64_t l[5];
    memcpy(l, in, sizeof l);
    uint64_t c0 = l[0] >> 51, c1 = l[1] >> 51, c2 = l[2] >> 51, c3 = l[3] >> 51, c4 = l[4] >> 51;
    l[0] &= MASK51; l[1] &= MASK51; l[2] &= MASK51; l[3] &= MASK51; l[4] &= MASK51;
    l[0] += c4 * 19; l[1] += c0; l[2] += c1; l[3] += c2; l[4] += c3;
    memcpy(o->v, l, sizeof l);
}

/* field.rs:58-73: limb-wise add, no reduction */
void fe_add(fe51 *o, const fe51 *a, const fe51 *b)
{
    for (int i = 0; i < 5; i++) o->v[i] = a->v[i] + b->v[i];
}

/* field.rs:82-102: add 16p, subtract, weak-reduce */
void fe_sub(fe51 *o, const fe51 *a, const fe51 *b)
{
    uint64_t t[5];
    t[0] = (a->v[0] + UINT64_C(36028797018963664)) - b->v[0];
    t[1] = (a->v[1] + UINT64_C(36028797018963952)) - b->v[1];
    t[2] = (a->v[2] + UINT64_C(36028797018963952)) - b->v[2];
    t[3] = (a->v[3] + UINT64_C(36028797018963952)) - b->v[3];
    t[4] = (a->v[4] + UINT64_C(36028797018963952)) - b->v[4];
    fe_reduce(o, t);
}

/* field.rs:276-287 */
void fe_neg(fe51 *o, const fe51 *a)
{
    uint64_t t[5];
    t[0] = UINT64_C(36028797018963664) - a->v[0];
    t[1] = UINT64_C(36028797018963952) - a->v[1];
    t[2] = UINT64_C(36028797018963952) - a->v[2];
    t[3] = UINT64_C(36028797018963952) - a->v[3];
    t[4] = UINT64_C(36028797018963952) - a->v[4];
    fe_reduce(o, t);
}

/* shared tail of mul / pow2k: field.rs:175-209 and :520-551 */
static void fe_carry_out(fe51 *o, u128 c0, u128 c1, u128 c2, u128 c3, u128 c4)
{
    uint64_t out[5];
    c1 += (uint64_t)(c0 >> 51); out[0] = (uint64_t)c0 & MASK51;
    c2 += (uint64_t)(c1 >> 51); out[1] = (uint64_t)c1 & MASK51;
    c3 += (uint64_t)(c2 >> 51); out[2] = (uint64_t)c2 & MASK51;
    c4 += (uint64_t)(c3 >> 51); out[3] = (uint64_t)c3 & MASK51;
    uint64_t carry = (uint64_t)(c4 >> 51); out[4] = (uint64_t)c4 & MASK51;
    out[0] += carry * 19;
    out[1] += out[0] >> 51;
    out[0] &= MASK51;
    memcpy(o->v, out, sizeof out);
}

/* field.rs:111-214 */
void fe_mul(fe51 *o, const fe51 *fa, const fe51 *fb)
{
    const uint64_t *a = fa->v, *b = fb->v;
    uint64_t b1_19 = b[1] * 19, b2_19 = b[2] * 19, b3_19 = b[3] * 19, b4_19 = b[4] * 19;
    u128 c0 = m(a[0], b[0]) + m(a[4], b1_19) + m(a[3], b2_19) + m(a[2], b3_19) + m(a[1], b4_19);
    u128 c1 = m(a[1], b[0]) + m(a[0], b[1]) + m(a[4], b2_19) + m(a[3], b3_19) + m(a[2], b4_19);
    u128 c2 = m(a[2], b[0]) + m(a[1], b[1]) + m(a[0], b[2]) + m(a[4], b3_19) + m(a[3], b4_19);
    u128 c3 = m(a[3], b[0]) + m(a[2], b[1]) + m(a[1], b[2]) + m(a[0], b[3]) + m(a[4], b4_19);
    u128 c4 = m(a[4], b[0]) + m(a[3], b[1]) + m(a[2], b[2]) + m(a[1], b[3]) + m(a[0], b[4]);
    fe_carry_out(o, c0, c1, c2, c3, c4);
}

/* field.rs:454-559 */
void fe_pow2k(fe51 *o, const fe51 *fa, uint32_t k)
{
    fe51 t = *fa;
    uint64_t *a = t.v;
    for (;;) {
        uint64_t a3_19 = 19 * a[3], a4_19 = 19 * a[4];
        u128 c0 = m(a[0], a[0]) + 2 * (m(a[1], a4_19) + m(a[2], a3_19));
        u128 c1 = m(a[3], a3_19) + 2 * (m(a[0], a[1]) + m(a[2], a4_19));
        u128 c2 = m(a[1], a[1]) + 2 * (m(a[0], a[2]) + m(a[4], a3_19));
        u128 c3 = m(a[4], a4_19) + 2 * (m(a[0], a[3]) + m(a[1], a[2]));
        u128 c4 = m(a[2], a[2]) + 2 * (m(a[0], a[4]) + m(a[1], a[3]));
        fe_carry_out(&t, c0, c1, c2, c3, c4);
        if (--k == 0) break;
    }
    *o = t;
}

void fe_square(fe51 *o, const fe51 *a) { fe_pow2k(o, a, 1); }

/* field.rs:567-574 */
void fe_square2(fe51 *o, const fe51 *a)
{
    fe_pow2k(o, a, 1);
    for (int i = 0; i < 5; i++) o->v[i] *= 2;
}

static uint64_t load8(const uint8_t *s)
{
    uint64_t r = 0;
    for (int i = 0; i < 8; i++) r |= (uint64_t)s[i] << (8 * i);
    return r;
}

/* field.rs:338-363: ignores bit 255, accepts non-canonical values */
void fe_from_bytes(fe51 *o, const uint8_t s[32])
{
    o->v[0] = load8(s) & MASK51;
    o->v[1] = (load8(s + 6) >> 3) & MASK51;
    o->v[2] = (load8(s + 12) >> 6) & MASK51;
    o->v[3] = (load8(s + 19) >> 1) & MASK51;
    o->v[4] = (load8(s + 24) >> 12) & MASK51;
}

/* field.rs:368-450: canonical encoding */
void fe_to_bytes(uint8_t s[32], const fe51 *a)
{
    fe51 r;
    fe_reduce(&r, a->v);
    uint64_t *l = r.v;
    uint64_t q = (l[0] + 19) >> 51;
    q = (l[1] + q) >> 51; q = (l[2] + q) >> 51; q = (l[3] + q) >> 51; q = (l[4] + q) >> 51;
    l[0] += 19 * q;
    l[1] += l[0] >> 51; l[0] &= MASK51;
    l[2] += l[1] >> 51; l[1] &= MASK51;
    l[3] += l[2] >> 51; l[2] &= MASK51;
    l[4] += l[3] >> 51; l[3] &= MASK51;
    l[4] &= MASK51;
    /* pack 5 x 51 bits little-endian */
    unsigned __int128 acc = 0; int bits = 0, k = 0;
    for (int i = 0; i < 5; i++) {
        acc |= (unsigned __int128)l[i] << bits; bits += 51;
        while (bits >= 8) { s[k++] = (uint8_t)acc; acc >>= 8; bits -= 8; }
    }
    s[k++] = (uint8_t)acc; /* 255 bits -> last byte has 7 bits */
}

/* C/field.rs:92-99: equality of canonical encodings */
int fe_ct_eq(const fe51 *a, const fe51 *b)
{
    uint8_t x[32], y[32];
    fe_to_bytes(x, a); fe_to_bytes(y, b);
    return memcmp(x, y, 32) == 0;
}

int fe_is_negative(const fe51 *a) { uint8_t x[32]; fe_to_bytes(x, a); return x[0] & 1; }

int fe_is_zero(const fe51 *a)
{
    uint8_t x[32]; fe_to_bytes(x, a);
    uint8_t acc = 0; for (int i = 0; i < 32; i++) acc |= x[i];
    return acc == 0;
}

void fe_cond_assign(fe51 *o, const fe51 *a, int c) { if (c) *o = *a; }
void fe_cond_negate(fe51 *o, int c) { if (c) { fe51 t; fe_neg(&t, o); *o = t; } }

/* C/field.rs:176-210: returns (a^(2^250-1), a^11) */
static void fe_pow22501(fe51 *t19, fe51 *t3, const fe51 *a)
{
    fe51 t0, t1, t2, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18;
    fe_square(&t0, a);
    fe_square(&t1, &t0); fe_square(&t1, &t1);
    fe_mul(&t2, a, &t1);
    fe_mul(t3, &t0, &t2);
    fe_square(&t4, t3);
    fe_mul(&t5, &t2, &t4);
    fe_pow2k(&t6, &t5, 5);
    fe_mul(&t7, &t6, &t5);
    fe_pow2k(&t8, &t7, 10);
    fe_mul(&t9, &t8, &t7);
    fe_pow2k(&t10, &t9, 20);
    fe_mul(&t11, &t10, &t9);
    fe_pow2k(&t12, &t11, 10);
    fe_mul(&t13, &t12, &t7);
    fe_pow2k(&t14, &t13, 50);
    fe_mul(&t15, &t14, &t13);
    fe_pow2k(&t16, &t15, 100);
    fe_mul(&t17, &t16, &t15);
    fe_pow2k(&t18, &t17, 50);
    fe_mul(t19, &t18, &t13);
}

/* C/field.rs:283-292 */
void fe_invert(fe51 *o, const fe51 *a)
{
    fe51 t19, t3, t20;
    fe_pow22501(&t19, &t3, a);
    fe_pow2k(&t20, &t19, 5);
    fe_mul(o, &t20, &t3);
}

/* C/field.rs:297-306 */
void fe_pow_p58(fe51 *o, const fe51 *a)
{
    fe51 t19, t3, t20;
    fe_pow22501(&t19, &t3, a);
    fe_pow2k(&t20, &t19, 2);
    fe_mul(o, a, &t20);
}

/* C/field.rs:320-366 */
int fe_sqrt_ratio_i(fe51 *out, const fe51 *u, const fe51 *v)
{
    fe51 v3, v7, r, t, check, i, neg_u, neg_u_i, r_prime;
    memcpy(i.v, K_SQRT_M1, sizeof i.v);
    fe_square(&t, v); fe_mul(&v3, &t, v);                 /* v3 = v^2 * v */
    fe_square(&t, &v3); fe_mul(&v7, &t, v);               /* v7 = v3^2 * v */
    fe51 uv3, uv7, pw;
    fe_mul(&uv3, u, &v3); fe_mul(&uv7, u, &v7);
    fe_pow_p58(&pw, &uv7);
    fe_mul(&r, &uv3, &pw);
    fe_square(&t, &r); fe_mul(&check, v, &t);

    fe_neg(&neg_u, u);
    fe_mul(&neg_u_i, &neg_u, &i);
    int correct_sign_sqrt = fe_ct_eq(&check, u);
    int flipped_sign_sqrt = fe_ct_eq(&check, &neg_u);
    int flipped_sign_sqrt_i = fe_ct_eq(&check, &neg_u_i);

    fe_mul(&r_prime, &i, &r);
    fe_cond_assign(&r, &r_prime, flipped_sign_sqrt | flipped_sign_sqrt_i);
    fe_cond_negate(&r, fe_is_negative(&r));
    *out = r;
    return correct_sign_sqrt | flipped_sign_sqrt;
}

int fe_invsqrt(fe51 *r, const fe51 *a) { fe51 one; fe_one(&one); return fe_sqrt_ratio_i(r, &one, a); }

/* C/field.rs:239-273: Montgomery's trick, zeros are skipped (left unchanged) */
void fe_invert_batch(fe51 *inputs, size_t n)
{
    fe51 *scratch = (fe51 *)malloc(sizeof(fe51) * (n ? n : 1));
    fe51 acc, tmp; fe_one(&acc);
    for (size_t i = 0; i < n; i++) {
        scratch[i] = acc;
        fe_mul(&tmp, &acc, &inputs[i]);
        fe_cond_assign(&acc, &tmp, !fe_is_zero(&inputs[i]));
    }
    fe_invert(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        fe51 t2;
        fe_mul(&tmp, &acc, &inputs[i]);
        int nz = !fe_is_zero(&inputs[i]);
        fe_mul(&t2, &acc, &scratch[i]);
        fe_cond_assign(&inputs[i], &t2, nz);
        fe_cond_assign(&acc, &tmp, nz);
    }
    free(scratch);
}
