/*
 * ristretto.c -- Ristretto255 encode/decode/equality and the double-base batch.
 * TEST INFRASTRUCTURE (oracle).  Restates C/ristretto.rs:266-345, :500-533,
 * :564-646, :815-830, :964-977.
 */
#include "oracle.h"
#include "constants.h"
#include <stdlib.h>
#include <string.h>

static void fe_const(fe51 *o, const uint64_t k[5]) { memcpy(o->v, k, sizeof o->v); }

/* C/ristretto.rs:266-345 */
int ristretto_decompress(ge_p3 *o, const uint8_t in[32])
{
    /* step_1 (:290-308) */
    fe51 s; uint8_t chk[32];
    fe_from_bytes(&s, in);
    fe_to_bytes(chk, &s);
    int canonical = memcmp(chk, in, 32) == 0;
    int s_neg = fe_is_negative(&s);
    if (!canonical || s_neg) return 0;

    /* step_2 (:310-345) */
    fe51 one, ss, u1, u2, u2_sqr, v, t, I, Dx, Dy, x, y, tt, nd, d;
    fe_one(&one); fe_const(&d, K_EDWARDS_D);
    fe_square(&ss, &s);
    fe_sub(&u1, &one, &ss);
    fe_add(&u2, &one, &ss);
    fe_square(&u2_sqr, &u2);
    fe_neg(&nd, &d);
    fe_square(&t, &u1); fe_mul(&t, &nd, &t); fe_sub(&v, &t, &u2_sqr);
    fe_mul(&t, &v, &u2_sqr);
    int ok = fe_invsqrt(&I, &t);
    fe_mul(&Dx, &I, &u2);
    fe_mul(&t, &Dx, &v); fe_mul(&Dy, &I, &t);
    fe_add(&t, &s, &s); fe_mul(&x, &t, &Dx);
    fe_cond_negate(&x, fe_is_negative(&x));
    fe_mul(&y, &u1, &Dy);
    fe_mul(&tt, &x, &y);
    if (!ok || fe_is_negative(&tt) || fe_is_zero(&y)) return 0;
    o->X = x; o->Y = y; o->Z = one; o->T = tt;
    return 1;
}

/* C/ristretto.rs:500-533 */
void ristretto_compress(uint8_t out[32], const ge_p3 *p)
{
    fe51 X = p->X, Y = p->Y; const fe51 *Z = &p->Z, *T = &p->T;
    fe51 u1, u2, t, t2, invsqrt, i1, i2, z_inv, den_inv, iX, iY, ench, sqrt_m1, magic, s;
    fe_const(&sqrt_m1, K_SQRT_M1); fe_const(&magic, K_INVSQRT_A_MINUS_D);
    fe_add(&t, Z, &Y); fe_sub(&t2, Z, &Y); fe_mul(&u1, &t, &t2);
    fe_mul(&u2, &X, &Y);
    fe_square(&t, &u2); fe_mul(&t, &u1, &t);
    (void)fe_invsqrt(&invsqrt, &t);
    fe_mul(&i1, &invsqrt, &u1);
    fe_mul(&i2, &invsqrt, &u2);
    fe_mul(&t, &i2, T); fe_mul(&z_inv, &i1, &t);
    den_inv = i2;
    fe_mul(&iX, &X, &sqrt_m1);
    fe_mul(&iY, &Y, &sqrt_m1);
    fe_mul(&ench, &i1, &magic);
    fe_mul(&t, T, &z_inv);
    int rotate = fe_is_negative(&t);
    fe_cond_assign(&X, &iY, rotate);
    fe_cond_assign(&Y, &iX, rotate);
    fe_cond_assign(&den_inv, &ench, rotate);
    fe_mul(&t, &X, &z_inv);
    fe_cond_negate(&Y, fe_is_negative(&t));
    fe_sub(&t, Z, &Y); fe_mul(&s, &den_inv, &t);
    fe_cond_negate(&s, fe_is_negative(&s));
    fe_to_bytes(out, &s);
}

/* C/ristretto.rs:815-830 */
int ristretto_ct_eq(const ge_p3 *a, const ge_p3 *b)
{
    fe51 x1y2, y1x2, x1x2, y1y2;
    fe_mul(&x1y2, &a->X, &b->Y); fe_mul(&y1x2, &a->Y, &b->X);
    fe_mul(&x1x2, &a->X, &b->X); fe_mul(&y1y2, &a->Y, &b->Y);
    return fe_ct_eq(&x1y2, &y1x2) | fe_ct_eq(&x1x2, &y1y2);
}

/* n independent RistrettoPoint::multiscalar_mul([a_i, b_i], [G, H]).compress()
 * (C/ristretto.rs:964-977 -> C/edwards.rs:970-995 -> straus.rs:103-144) */
int ristretto_double_base_batch(uint8_t *out, const uint8_t *a, const uint8_t *b,
                                const uint8_t G[32], const uint8_t H[32], size_t n)
{
    ge_p3 pts[2];
    if (!ristretto_decompress(&pts[0], G) || !ristretto_decompress(&pts[1], H)) return 1;
    for (size_t i = 0; i < n; i++) {
        uint8_t sc[64]; ge_p3 r;
        memcpy(sc, a + 32 * i, 32); memcpy(sc + 32, b + 32 * i, 32);
        edwards_multiscalar_mul(&r, sc, pts, 2);
        ristretto_compress(out + 32 * i, &r);
    }
    return 0;
}

/* RistrettoPoint::double_and_compress_batch, C/ristretto.rs:564-646: out[i] = compress(2 P_i) with one
 * simultaneous inversion (C/field.rs:239-273) and no square roots */
void ristretto_double_and_compress_batch(uint8_t *out, const ge_p3 *points, size_t n)
{
    typedef struct { fe51 e, f, g, h, eg, fh; } state;
    state *st = (state *)malloc(sizeof(state) * (n ? n : 1));
    fe51 *invs = (fe51 *)malloc(sizeof(fe51) * (n ? n : 1));
    fe51 d, sqrt_m1, invsqrt_a_minus_d;
    fe_const(&d, K_EDWARDS_D); fe_const(&sqrt_m1, K_SQRT_M1); fe_const(&invsqrt_a_minus_d, K_INVSQRT_A_MINUS_D);
    for (size_t i = 0; i < n; i++) {                              /* :584-601 */
        const ge_p3 *P = &points[i];
        fe51 XX, YY, ZZ, dTT, t;
        fe_square(&XX, &P->X); fe_square(&YY, &P->Y); fe_square(&ZZ, &P->Z);
        fe_square(&t, &P->T); fe_mul(&dTT, &t, &d);
        fe_add(&t, &P->Y, &P->Y); fe_mul(&st[i].e, &P->X, &t);
        fe_add(&st[i].f, &ZZ, &dTT);
        fe_add(&st[i].g, &YY, &XX);
        fe_sub(&st[i].h, &ZZ, &dTT);
        fe_mul(&st[i].eg, &st[i].e, &st[i].g);
        fe_mul(&st[i].fh, &st[i].f, &st[i].h);
        fe_mul(&invs[i], &st[i].eg, &st[i].fh);                   /* :607 */
    }
    fe_invert_batch(invs, n);                                     /* :609 */
    for (size_t i = 0; i < n; i++) {                              /* :611-644 */
        fe51 Zinv, Tinv, magic = invsqrt_a_minus_d, t, e = st[i].e, g = st[i].g, h = st[i].h, minus_e, f_times_sqrta, s;
        fe_mul(&Zinv, &st[i].eg, &invs[i]);
        fe_mul(&Tinv, &st[i].fh, &invs[i]);
        fe_mul(&t, &st[i].eg, &Zinv);
        int negcheck1 = fe_is_negative(&t);
        fe_neg(&minus_e, &e);
        fe_mul(&f_times_sqrta, &st[i].f, &sqrt_m1);
        fe_cond_assign(&e, &st[i].g, negcheck1);
        fe_cond_assign(&g, &minus_e, negcheck1);
        fe_cond_assign(&h, &f_times_sqrta, negcheck1);
        fe_cond_assign(&magic, &sqrt_m1, negcheck1);
        fe_mul(&t, &h, &e); fe_mul(&t, &t, &Zinv);
        int negcheck2 = fe_is_negative(&t);
        fe_cond_negate(&g, negcheck2);
        fe51 hg, gt;
        fe_sub(&hg, &h, &g);
        fe_mul(&gt, &g, &Tinv); fe_mul(&gt, &magic, &gt);
        fe_mul(&s, &hg, &gt);
        fe_cond_negate(&s, fe_is_negative(&s));
        fe_to_bytes(out + 32 * i, &s);
    }
    free(st); free(invs);
}
