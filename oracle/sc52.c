/*
 * sc52.c -- arithmetic mod l = 2^252 + 27742317777372353535851937790883648493
 * in radix 2^52 (Montgomery, R = 2^260) and the Scalar digit recoders.
 * TEST INFRASTRUCTURE (oracle).  Restates C/backend/serial/u64/scalar.rs and
 * the parts of C/scalar.rs on the multiscalar / verify_batch path.
 */
#include "oracle.h"
#include "constants.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define MASK52 ((UINT64_C(1) << 52) - 1)
static inline u128 m(uint64_t x, uint64_t y) { return (u128)x * (u128)y; }

static void load_words(uint64_t *w, const uint8_t *s, int nwords)
{
    for (int i = 0; i < nwords; i++) {
        w[i] = 0;
        for (int j = 0; j < 8; j++) w[i] |= (uint64_t)s[i * 8 + j] << (8 * j);
    }
}

/* u64/scalar.rs:66-86 */
void sc52_from_bytes(sc52 *o, const uint8_t s[32])
{
    uint64_t w[4];
    load_words(w, s, 4);
    uint64_t top_mask = (UINT64_C(1) << 48) - 1;
    o->v[0] = w[0] & MASK52;
    o->v[1] = ((w[0] >> 52) | (w[1] << 12)) & MASK52;
    o->v[2] = ((w[1] >> 40) | (w[2] << 24)) & MASK52;
    o->v[3] = ((w[2] >> 28) | (w[3] << 36)) & MASK52;
    o->v[4] = (w[3] >> 16) & top_mask;
}

/* u64/scalar.rs:121-158 */
void sc52_to_bytes(uint8_t s[32], const sc52 *a)
{
    u128 acc = 0; int bits = 0, k = 0;
    for (int i = 0; i < 5; i++) {
        acc |= (u128)a->v[i] << bits; bits += 52;
        while (bits >= 8 && k < 32) { s[k++] = (uint8_t)acc; acc >>= 8; bits -= 8; }
    }
}

/* u64/scalar.rs:194-206 */
static uint64_t sc52_conditional_add_l(sc52 *a, int cond)
{
    uint64_t carry = 0;
    for (int i = 0; i < 5; i++) {
        uint64_t addend = cond ? K_SC_L[i] : 0;
        carry = (carry >> 52) + a->v[i] + addend;
        a->v[i] = carry & MASK52;
    }
    return carry;
}

/* u64/scalar.rs:177-191 */
void sc52_sub(sc52 *o, const sc52 *a, const sc52 *b)
{
    sc52 d;
    uint64_t borrow = 0;
    for (int i = 0; i < 5; i++) {
        borrow = a->v[i] - (b->v[i] + (borrow >> 63));
        d.v[i] = borrow & MASK52;
    }
    sc52_conditional_add_l(&d, (int)(borrow >> 63));
    *o = d;
}

/* u64/scalar.rs:161-174 */
void sc52_add(sc52 *o, const sc52 *a, const sc52 *b)
{
    sc52 sum, l;
    uint64_t carry = 0;
    for (int i = 0; i < 5; i++) {
        carry = a->v[i] + b->v[i] + (carry >> 52);
        sum.v[i] = carry & MASK52;
    }
    memcpy(l.v, K_SC_L, sizeof l.v);
    sc52_sub(o, &sum, &l);
}

/* u64/scalar.rs:222-236 */
static void sc52_mul_internal(u128 z[9], const sc52 *sa, const sc52 *sb)
{
    const uint64_t *a = sa->v, *b = sb->v;
    z[0] = m(a[0], b[0]);
    z[1] = m(a[0], b[1]) + m(a[1], b[0]);
    z[2] = m(a[0], b[2]) + m(a[1], b[1]) + m(a[2], b[0]);
    z[3] = m(a[0], b[3]) + m(a[1], b[2]) + m(a[2], b[1]) + m(a[3], b[0]);
    z[4] = m(a[0], b[4]) + m(a[1], b[3]) + m(a[2], b[2]) + m(a[3], b[1]) + m(a[4], b[0]);
    z[5] = m(a[1], b[4]) + m(a[2], b[3]) + m(a[3], b[2]) + m(a[4], b[1]);
    z[6] = m(a[2], b[4]) + m(a[3], b[3]) + m(a[4], b[2]);
    z[7] = m(a[3], b[4]) + m(a[4], b[3]);
    z[8] = m(a[4], b[4]);
}

/* u64/scalar.rs:265-298 */
static void sc52_montgomery_reduce(sc52 *o, const u128 limbs[9])
{
    const uint64_t *l = K_SC_L;
    u128 carry, sum; uint64_t n0, n1, n2, n3, n4, r0, r1, r2, r3, r4;
#define PART1(S, N) do { sum = (S); N = ((uint64_t)sum * K_SC_LFACTOR) & MASK52; \
                         carry = (sum + m(N, l[0])) >> 52; } while (0)
#define PART2(S, R) do { sum = (S); R = (uint64_t)sum & MASK52; carry = sum >> 52; } while (0)
    PART1(limbs[0], n0);
    PART1(carry + limbs[1] + m(n0, l[1]), n1);
    PART1(carry + limbs[2] + m(n0, l[2]) + m(n1, l[1]), n2);
    PART1(carry + limbs[3] + m(n1, l[2]) + m(n2, l[1]), n3);
    PART1(carry + limbs[4] + m(n0, l[4]) + m(n2, l[2]) + m(n3, l[1]), n4);
    PART2(carry + limbs[5] + m(n1, l[4]) + m(n3, l[2]) + m(n4, l[1]), r0);
    PART2(carry + limbs[6] + m(n2, l[4]) + m(n4, l[2]), r1);
    PART2(carry + limbs[7] + m(n3, l[4]), r2);
    PART2(carry + limbs[8] + m(n4, l[4]), r3);
    r4 = (uint64_t)carry;
#undef PART1
#undef PART2
    sc52 r = {{r0, r1, r2, r3, r4}}, ll;
    memcpy(ll.v, l, sizeof ll.v);
    sc52_sub(o, &r, &ll);
}

/* u64/scalar.rs:317-319 */
void sc52_montgomery_mul(sc52 *o, const sc52 *a, const sc52 *b)
{
    u128 z[9];
    sc52_mul_internal(z, a, b);
    sc52_montgomery_reduce(o, z);
}

/* u64/scalar.rs:302-305 */
void sc52_mul(sc52 *o, const sc52 *a, const sc52 *b)
{
    sc52 ab, rr;
    memcpy(rr.v, K_SC_RR, sizeof rr.v);
    sc52_montgomery_mul(&ab, a, b);
    sc52_montgomery_mul(o, &ab, &rr);
}

/* u64/scalar.rs:89-116 */
void sc52_from_bytes_wide(sc52 *o, const uint8_t s[64])
{
    uint64_t w[8];
    load_words(w, s, 8);
    sc52 lo, hi, r, rr;
    lo.v[0] = w[0] & MASK52;
    lo.v[1] = ((w[0] >> 52) | (w[1] << 12)) & MASK52;
    lo.v[2] = ((w[1] >> 40) | (w[2] << 24)) & MASK52;
    lo.v[3] = ((w[2] >> 28) | (w[3] << 36)) & MASK52;
    lo.v[4] = ((w[3] >> 16) | (w[4] << 48)) & MASK52;
    hi.v[0] = (w[4] >> 4) & MASK52;
    hi.v[1] = ((w[4] >> 56) | (w[5] << 8)) & MASK52;
    hi.v[2] = ((w[5] >> 44) | (w[6] << 20)) & MASK52;
    hi.v[3] = ((w[6] >> 32) | (w[7] << 32)) & MASK52;
    hi.v[4] = w[7] >> 20;
    memcpy(r.v, K_SC_R, sizeof r.v);
    memcpy(rr.v, K_SC_RR, sizeof rr.v);
    sc52_montgomery_mul(&lo, &lo, &r);
    sc52_montgomery_mul(&hi, &hi, &rr);
    sc52_add(o, &hi, &lo);
}

/* ---- Scalar (32 bytes) level, C/scalar.rs ---- */

/* C/scalar.rs:1159-1165 `reduce` */
void scalar_reduce(uint8_t o[32], const uint8_t a[32])
{
    sc52 x, r, xm;
    u128 z[9];
    sc52_from_bytes(&x, a);
    memcpy(r.v, K_SC_R, sizeof r.v);
    sc52_mul_internal(z, &x, &r);
    sc52_montgomery_reduce(&xm, z);
    sc52_to_bytes(o, &xm);
}

void scalar_from_bytes_mod_order_wide(uint8_t o[32], const uint8_t a[64])
{
    sc52 x;
    sc52_from_bytes_wide(&x, a);
    sc52_to_bytes(o, &x);
}

/* C/scalar.rs:259-263 + :1168-1170 (from_canonical_bytes: high bit unset AND s == s.reduce()) */
int scalar_is_canonical(const uint8_t a[32])
{
    uint8_t r[32];
    scalar_reduce(r, a);
    return ((a[31] >> 7) == 0) && memcmp(r, a, 32) == 0;
}

void scalar_add(uint8_t o[32], const uint8_t a[32], const uint8_t b[32])
{
    sc52 x, y, z;
    sc52_from_bytes(&x, a); sc52_from_bytes(&y, b);
    sc52_add(&z, &x, &y);
    sc52_to_bytes(o, &z);
}

void scalar_sub(uint8_t o[32], const uint8_t a[32], const uint8_t b[32])
{
    sc52 x, y, z;
    sc52_from_bytes(&x, a); sc52_from_bytes(&y, b);
    sc52_sub(&z, &x, &y);
    sc52_to_bytes(o, &z);
}

void scalar_mul(uint8_t o[32], const uint8_t a[32], const uint8_t b[32])
{
    sc52 x, y, z;
    sc52_from_bytes(&x, a); sc52_from_bytes(&y, b);
    sc52_mul(&z, &x, &y);
    sc52_to_bytes(o, &z);
}

/* C/scalar.rs:366-374: reduce first, then 0 - x */
void scalar_neg(uint8_t o[32], const uint8_t a[32])
{
    sc52 x, r, xm, zero, z;
    u128 w[9];
    sc52_from_bytes(&x, a);
    memcpy(r.v, K_SC_R, sizeof r.v);
    sc52_mul_internal(w, &x, &r);
    sc52_montgomery_reduce(&xm, w);
    memset(&zero, 0, sizeof zero);
    sc52_sub(&z, &zero, &xm);
    sc52_to_bytes(o, &z);
}

void scalar_from_u64(uint8_t o[32], uint64_t x)
{
    memset(o, 0, 32);
    for (int i = 0; i < 8; i++) o[i] = (uint8_t)(x >> (8 * i));
}

/* value of C/scalar.rs:739 `invert` (a^(l-2)); plain square-and-multiply, the reference's
 * addition chain (:1240) yields the same canonical value */
void scalar_invert(uint8_t o[32], const uint8_t a[32])
{
    /* l - 2 little-endian */
    uint8_t e[32];
    sc52 lm; memcpy(lm.v, K_SC_L, sizeof lm.v);
    sc52_to_bytes(e, &lm);
    e[0] -= 2; /* l ends in ...ed, no borrow */
    uint8_t acc[32], base[32];
    scalar_from_u64(acc, 1);
    scalar_reduce(base, a);
    for (int i = 252; i >= 0; i--) {
        scalar_mul(acc, acc, acc);
        if ((e[i >> 3] >> (i & 7)) & 1) scalar_mul(acc, acc, base);
    }
    memcpy(o, acc, 32);
}

/* C/scalar.rs:955-1007 */
void scalar_non_adjacent_form(int8_t naf[256], const uint8_t a[32], unsigned w)
{
    memset(naf, 0, 256);
    uint64_t x[5] = {0, 0, 0, 0, 0};
    load_words(x, a, 4);
    uint64_t width = UINT64_C(1) << w, window_mask = width - 1;
    size_t pos = 0; uint64_t carry = 0;
    while (pos < 256) {
        size_t idx = pos / 64, bit = pos % 64;
        uint64_t bit_buf;
        if (bit < 64 - w) bit_buf = x[idx] >> bit;
        else bit_buf = (x[idx] >> bit) | (x[1 + idx] << (64 - bit));
        uint64_t window = carry + (bit_buf & window_mask);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)((int8_t)window - (int8_t)width); }
        pos += w;
    }
}

/* C/scalar.rs:1019-1051 */
void scalar_as_radix_16(int8_t out[64], const uint8_t a[32])
{
    for (int i = 0; i < 32; i++) {
        out[2 * i] = (int8_t)(a[i] & 15);
        out[2 * i + 1] = (int8_t)((a[i] >> 4) & 15);
    }
    for (int i = 0; i < 63; i++) {
        int8_t carry = (int8_t)((out[i] + 8) >> 4);
        out[i] = (int8_t)(out[i] - (carry << 4));
        out[i + 1] = (int8_t)(out[i + 1] + carry);
    }
}

/* C/scalar.rs:1056-1069 */
size_t scalar_to_radix_2w_size_hint(unsigned w)
{
    size_t c = (256 + w - 1) / w;
    return w == 8 ? c + 1 : c;
}

/* C/scalar.rs:1093-1150 */
void scalar_as_radix_2w(int8_t digits[64], const uint8_t a[32], unsigned w)
{
    if (w == 4) { scalar_as_radix_16(digits, a); return; }
    memset(digits, 0, 64);
    uint64_t s[4];
    load_words(s, a, 4);
    uint64_t radix = UINT64_C(1) << w, window_mask = radix - 1, carry = 0;
    size_t digits_count = (256 + w - 1) / w;
    for (size_t i = 0; i < digits_count; i++) {
        size_t bit_offset = i * w, idx = bit_offset / 64, bit = bit_offset % 64;
        uint64_t bit_buf;
        if (bit < 64 - w || idx == 3) bit_buf = s[idx] >> bit;
        else bit_buf = (s[idx] >> bit) | (s[1 + idx] << (64 - bit));
        uint64_t coef = carry + (bit_buf & window_mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int8_t)((int64_t)coef - (int64_t)(carry << w));
    }
    if (w == 8) digits[digits_count] = (int8_t)(digits[digits_count] + (int8_t)carry);
    else digits[digits_count - 1] = (int8_t)(digits[digits_count - 1] + (int8_t)(carry << w));
}

/* Scalar::invert_batch_alloc (C/scalar.rs:793-853), values only: every scalar replaced by its inverse mod l,
 * ret = the product of all inverses.  Inputs must be nonzero (scalar.rs:796-799). */
void scalar_invert_batch(uint8_t *inout, size_t n, uint8_t ret[32])
{
    uint8_t *scratch = (uint8_t *)malloc(32 * (n ? n : 1));
    uint8_t acc[32], tmp[32];
    scalar_from_u64(acc, 1);
    for (size_t i = 0; i < n; i++) {                              /* :819-828 */
        memcpy(scratch + 32 * i, acc, 32);
        scalar_reduce(inout + 32 * i, inout + 32 * i);
        scalar_mul(acc, acc, inout + 32 * i);
    }
    scalar_invert(acc, acc);                                      /* :834 */
    memcpy(ret, acc, 32);                                         /* :837 */
    for (size_t i = n; i-- > 0;) {                                /* :841-846 */
        scalar_mul(tmp, acc, inout + 32 * i);
        scalar_mul(inout + 32 * i, acc, scratch + 32 * i);
        memcpy(acc, tmp, 32);
    }
    free(scratch);
}
