/*
 * curve.c -- twisted Edwards curve models, window tables, (de)compression.
 * TEST INFRASTRUCTURE (oracle).  Restates C/backend/serial/curve_models.rs,
 * C/window.rs and the hot-path parts of C/edwards.rs.
 */
#include "oracle.h"
#include "constants.h"
#include <string.h>

static void fe_const(fe51 *o, const uint64_t k[5]) { memcpy(o->v, k, sizeof o->v); }

/* C/edwards.rs:428-437 */
void ge_identity(ge_p3 *o) { fe_zero(&o->X); fe_one(&o->Y); fe_one(&o->Z); fe_zero(&o->T); }
void ge_p2_identity(ge_p2 *o) { fe_zero(&o->X); fe_one(&o->Y); fe_one(&o->Z); }

/* u64/constants.rs:163-186 (values derived in gen_constants.py) */
void ge_basepoint(ge_p3 *o)
{
    fe_const(&o->X, K_BASE_X); fe_const(&o->Y, K_BASE_Y); fe_one(&o->Z); fe_const(&o->T, K_BASE_T);
}

void ge_p3_from_limbs(ge_p3 *o, const uint64_t l[20])
{
    memcpy(o->X.v, l, 40); memcpy(o->Y.v, l + 5, 40); memcpy(o->Z.v, l + 10, 40); memcpy(o->T.v, l + 15, 40);
}
void ge_p3_to_limbs(uint64_t l[20], const ge_p3 *p)
{
    memcpy(l, p->X.v, 40); memcpy(l + 5, p->Y.v, 40); memcpy(l + 10, p->Z.v, 40); memcpy(l + 15, p->T.v, 40);
}

/* curve_models.rs:338-345: 3M + 1S */
void ge_p2_to_p3(ge_p3 *o, const ge_p2 *p)
{
    ge_p3 r;
    fe_mul(&r.X, &p->X, &p->Z); fe_mul(&r.Y, &p->Y, &p->Z); fe_square(&r.Z, &p->Z); fe_mul(&r.T, &p->X, &p->Y);
    *o = r;
}

/* curve_models.rs:353-359: 3M */
void ge_p1p1_to_p2(ge_p2 *o, const ge_p1p1 *p)
{
    ge_p2 r;
    fe_mul(&r.X, &p->X, &p->T); fe_mul(&r.Y, &p->Y, &p->Z); fe_mul(&r.Z, &p->Z, &p->T);
    *o = r;
}

/* curve_models.rs:365-372: 4M */
void ge_p1p1_to_p3(ge_p3 *o, const ge_p1p1 *p)
{
    ge_p3 r;
    fe_mul(&r.X, &p->X, &p->T); fe_mul(&r.Y, &p->Y, &p->Z); fe_mul(&r.Z, &p->Z, &p->T); fe_mul(&r.T, &p->X, &p->Y);
    *o = r;
}

/* curve_models.rs:381-397: 4S */
void ge_p2_double(ge_p1p1 *o, const ge_p2 *p)
{
    fe51 XX, YY, ZZ2, XpY, XpY2, YYpXX, YYmXX;
    fe_square(&XX, &p->X); fe_square(&YY, &p->Y); fe_square2(&ZZ2, &p->Z);
    fe_add(&XpY, &p->X, &p->Y); fe_square(&XpY2, &XpY);
    fe_add(&YYpXX, &YY, &XX); fe_sub(&YYmXX, &YY, &XX);
    fe_sub(&o->X, &XpY2, &YYpXX);
    o->Y = YYpXX;
    o->Z = YYmXX;
    fe_sub(&o->T, &ZZ2, &YYmXX);
}

/* curve_models.rs:411-430 */
void ge_add_pniels(ge_p1p1 *o, const ge_p3 *p, const ge_pniels *q)
{
    fe51 YpX, YmX, PP, MM, TT2d, ZZ, ZZ2;
    fe_add(&YpX, &p->Y, &p->X); fe_sub(&YmX, &p->Y, &p->X);
    fe_mul(&PP, &YpX, &q->Y_plus_X); fe_mul(&MM, &YmX, &q->Y_minus_X);
    fe_mul(&TT2d, &p->T, &q->T2d); fe_mul(&ZZ, &p->Z, &q->Z);
    fe_add(&ZZ2, &ZZ, &ZZ);
    fe_sub(&o->X, &PP, &MM); fe_add(&o->Y, &PP, &MM);
    fe_add(&o->Z, &ZZ2, &TT2d); fe_sub(&o->T, &ZZ2, &TT2d);
}

/* curve_models.rs:433-452 */
void ge_sub_pniels(ge_p1p1 *o, const ge_p3 *p, const ge_pniels *q)
{
    fe51 YpX, YmX, PM, MP, TT2d, ZZ, ZZ2;
    fe_add(&YpX, &p->Y, &p->X); fe_sub(&YmX, &p->Y, &p->X);
    fe_mul(&PM, &YpX, &q->Y_minus_X); fe_mul(&MP, &YmX, &q->Y_plus_X);
    fe_mul(&TT2d, &p->T, &q->T2d); fe_mul(&ZZ, &p->Z, &q->Z);
    fe_add(&ZZ2, &ZZ, &ZZ);
    fe_sub(&o->X, &PM, &MP); fe_add(&o->Y, &PM, &MP);
    fe_sub(&o->Z, &ZZ2, &TT2d); fe_add(&o->T, &ZZ2, &TT2d);
}

/* curve_models.rs:455-473 */
void ge_add_aniels(ge_p1p1 *o, const ge_p3 *p, const ge_aniels *q)
{
    fe51 YpX, YmX, PP, MM, Txy2d, Z2;
    fe_add(&YpX, &p->Y, &p->X); fe_sub(&YmX, &p->Y, &p->X);
    fe_mul(&PP, &YpX, &q->y_plus_x); fe_mul(&MM, &YmX, &q->y_minus_x);
    fe_mul(&Txy2d, &p->T, &q->xy2d); fe_add(&Z2, &p->Z, &p->Z);
    fe_sub(&o->X, &PP, &MM); fe_add(&o->Y, &PP, &MM);
    fe_add(&o->Z, &Z2, &Txy2d); fe_sub(&o->T, &Z2, &Txy2d);
}

/* curve_models.rs:476-494 */
void ge_sub_aniels(ge_p1p1 *o, const ge_p3 *p, const ge_aniels *q)
{
    fe51 YpX, YmX, PM, MP, Txy2d, Z2;
    fe_add(&YpX, &p->Y, &p->X); fe_sub(&YmX, &p->Y, &p->X);
    fe_mul(&PM, &YpX, &q->y_minus_x); fe_mul(&MP, &YmX, &q->y_plus_x);
    fe_mul(&Txy2d, &p->T, &q->xy2d); fe_add(&Z2, &p->Z, &p->Z);
    fe_sub(&o->X, &PM, &MP); fe_add(&o->Y, &PM, &MP);
    fe_sub(&o->Z, &Z2, &Txy2d); fe_add(&o->T, &Z2, &Txy2d);
}

/* curve_models.rs:500-511 */
void ge_pniels_neg(ge_pniels *o, const ge_pniels *p)
{
    ge_pniels r;
    r.Y_plus_X = p->Y_minus_X; r.Y_minus_X = p->Y_plus_X; r.Z = p->Z; fe_neg(&r.T2d, &p->T2d);
    *o = r;
}

/* C/edwards.rs:528-535 */
void ge_p3_to_pniels(ge_pniels *o, const ge_p3 *p)
{
    fe51 d2; fe_const(&d2, K_EDWARDS_D2);
    ge_pniels r;
    fe_add(&r.Y_plus_X, &p->Y, &p->X); fe_sub(&r.Y_minus_X, &p->Y, &p->X);
    r.Z = p->Z; fe_mul(&r.T2d, &p->T, &d2);
    *o = r;
}

void ge_p3_to_p2(ge_p2 *o, const ge_p3 *p) { o->X = p->X; o->Y = p->Y; o->Z = p->Z; }

/* C/edwards.rs:786-788 */
void ge_p3_double(ge_p3 *o, const ge_p3 *p)
{
    ge_p2 s; ge_p1p1 r;
    ge_p3_to_p2(&s, p); ge_p2_double(&r, &s); ge_p1p1_to_p3(o, &r);
}

/* C/edwards.rs:795-800 */
void ge_p3_add(ge_p3 *o, const ge_p3 *p, const ge_p3 *q)
{
    ge_pniels n; ge_p1p1 r;
    ge_p3_to_pniels(&n, q); ge_add_pniels(&r, p, &n); ge_p1p1_to_p3(o, &r);
}

/* C/edwards.rs:818-823 */
void ge_p3_sub(ge_p3 *o, const ge_p3 *p, const ge_p3 *q)
{
    ge_pniels n; ge_p1p1 r;
    ge_p3_to_pniels(&n, q); ge_sub_pniels(&r, p, &n); ge_p1p1_to_p3(o, &r);
}

/* C/edwards.rs:853-864 */
void ge_p3_neg(ge_p3 *o, const ge_p3 *p)
{
    ge_p3 r;
    fe_neg(&r.X, &p->X); r.Y = p->Y; r.Z = p->Z; fe_neg(&r.T, &p->T);
    *o = r;
}

/* C/edwards.rs:1370-1380 */
void ge_mul_by_pow_2(ge_p3 *o, const ge_p3 *p, uint32_t k)
{
    ge_p1p1 r; ge_p2 s;
    ge_p3_to_p2(&s, p);
    for (uint32_t i = 0; i + 1 < k; i++) { ge_p2_double(&r, &s); ge_p1p1_to_p2(&s, &r); }
    ge_p2_double(&r, &s);
    ge_p1p1_to_p3(o, &r);
}

/* C/edwards.rs:501-512 */
int ge_p3_ct_eq(const ge_p3 *a, const ge_p3 *b)
{
    fe51 l, r; int ok;
    fe_mul(&l, &a->X, &b->Z); fe_mul(&r, &b->X, &a->Z); ok = fe_ct_eq(&l, &r);
    fe_mul(&l, &a->Y, &b->Z); fe_mul(&r, &b->Y, &a->Z); ok &= fe_ct_eq(&l, &r);
    return ok;
}

/* C/traits.rs:41-48 */
int ge_is_identity(const ge_p3 *a) { ge_p3 id; ge_identity(&id); return ge_p3_ct_eq(a, &id); }

/* C/edwards.rs:211-257 */
int ge_decompress(ge_p3 *o, const uint8_t s[32])
{
    fe51 Y, Z, YY, u, v, X, d;
    fe_const(&d, K_EDWARDS_D);
    fe_from_bytes(&Y, s); fe_one(&Z);
    fe_square(&YY, &Y);
    fe_sub(&u, &YY, &Z);
    fe_mul(&v, &YY, &d); fe_add(&v, &v, &Z);
    int ok = fe_sqrt_ratio_i(&X, &u, &v);
    if (!ok) return 0;
    fe_cond_negate(&X, s[31] >> 7);
    o->X = X; o->Y = Y; o->Z = Z; fe_mul(&o->T, &X, &Y);
    return 1;
}

/* C/edwards.rs:564-574 to_affine + C/edwards/affine.rs:71-75 compress */
void ge_compress(uint8_t s[32], const ge_p3 *p)
{
    fe51 recip, x, y;
    fe_invert(&recip, &p->Z);
    fe_mul(&x, &p->X, &recip); fe_mul(&y, &p->Y, &recip);
    fe_to_bytes(s, &y);
    s[31] ^= (uint8_t)(fe_is_negative(&x) << 7);
}

/* C/window.rs:97-105: [P, 2P, ..., 8P] */
void ge_lookup_table_from(ge_lookup_table *t, const ge_p3 *p)
{
    ge_p3_to_pniels(&t->t[0], p);
    for (int j = 0; j < 7; j++) {
        ge_p1p1 r; ge_p3 e;
        ge_add_pniels(&r, p, &t->t[j]); ge_p1p1_to_p3(&e, &r); ge_p3_to_pniels(&t->t[j + 1], &e);
    }
}

static void pniels_identity(ge_pniels *o)
{
    fe_one(&o->Y_plus_X); fe_one(&o->Y_minus_X); fe_one(&o->Z); fe_zero(&o->T2d);
}

/* C/window.rs:54-76: masked scan over all 8 entries + conditional negate */
void ge_lookup_table_select(ge_pniels *o, const ge_lookup_table *t, int8_t x)
{
    int16_t xmask = (int16_t)x >> 7;
    int16_t xabs = (int16_t)(((int16_t)x + xmask) ^ xmask);
    ge_pniels r; pniels_identity(&r);
    for (int j = 1; j < 9; j++) {
        int c = ((uint16_t)xabs == (uint16_t)j);
        fe_cond_assign(&r.Y_plus_X, &t->t[j - 1].Y_plus_X, c);
        fe_cond_assign(&r.Y_minus_X, &t->t[j - 1].Y_minus_X, c);
        fe_cond_assign(&r.Z, &t->t[j - 1].Z, c);
        fe_cond_assign(&r.T2d, &t->t[j - 1].T2d, c);
    }
    if (xmask & 1) { ge_pniels n; ge_pniels_neg(&n, &r); r = n; }
    *o = r;
}

/* C/window.rs:201-211: [A, 3A, ..., 15A] */
void ge_naf_table5_from(ge_naf_table5 *t, const ge_p3 *a)
{
    ge_p3 a2;
    ge_p3_to_pniels(&t->t[0], a);
    ge_p3_double(&a2, a);
    for (int i = 0; i < 7; i++) {
        ge_p1p1 r; ge_p3 e;
        ge_add_pniels(&r, &a2, &t->t[i]); ge_p1p1_to_p3(&e, &r); ge_p3_to_pniels(&t->t[i + 1], &e);
    }
}

/* C/backend/serial/scalar_mul/variable_base.rs:11-48 */
void ge_scalarmul(ge_p3 *o, const uint8_t scalar[32], const ge_p3 *p)
{
    ge_lookup_table tab; int8_t digits[64];
    ge_pniels sel; ge_p1p1 tmp1; ge_p2 tmp2; ge_p3 tmp3;
    ge_lookup_table_from(&tab, p);
    scalar_as_radix_16(digits, scalar);
    ge_identity(&tmp3);
    ge_lookup_table_select(&sel, &tab, digits[63]);
    ge_add_pniels(&tmp1, &tmp3, &sel);
    for (int i = 62; i >= 0; i--) {
        ge_p1p1_to_p2(&tmp2, &tmp1); ge_p2_double(&tmp1, &tmp2);
        ge_p1p1_to_p2(&tmp2, &tmp1); ge_p2_double(&tmp1, &tmp2);
        ge_p1p1_to_p2(&tmp2, &tmp1); ge_p2_double(&tmp1, &tmp2);
        ge_p1p1_to_p2(&tmp2, &tmp1); ge_p2_double(&tmp1, &tmp2);
        ge_p1p1_to_p3(&tmp3, &tmp1);
        ge_lookup_table_select(&sel, &tab, digits[i]);
        ge_add_pniels(&tmp1, &tmp3, &sel);
    }
    ge_p1p1_to_p3(o, &tmp1);
}

/* C/edwards.rs:1405-1407 */
int ge_is_small_order(const ge_p3 *p) { ge_p3 r; ge_mul_by_pow_2(&r, p, 3); return ge_is_identity(&r); }

/* C/edwards.rs:1435-1437: [l]P == identity.  l does not satisfy Scalar invariant #1's use in
 * as_radix_16 only through its top bit (l < 2^253), so the radix-16 ladder applies. */
int ge_is_torsion_free(const ge_p3 *p)
{
    uint8_t l[32]; sc52 lm; memcpy(lm.v, K_SC_L, sizeof lm.v); sc52_to_bytes(l, &lm);
    ge_p3 r; ge_scalarmul(&r, l, p);
    return ge_is_identity(&r);
}
