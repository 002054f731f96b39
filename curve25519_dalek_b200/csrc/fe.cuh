// fe.cuh -- GF(2^255-19) for sm_100a: ten 32-bit limbs in radix 2^25.5, every product a single
// IMAD.WIDE.U32 (32x32->64 multiply-add into a 64-bit column accumulator held in a register
// pair), no inter-product carries.
//
// The reference computes in radix 2^51 with u64 x u64 -> u128 products
// (curve25519-dalek/src/backend/serial/u64/field.rs:111-214, :454-559).  The SM has no 64-bit
// multiplier, so a 51-bit limb product would cost four IMAD.WIDE plus carry glue; 10 x 25.5-bit
// limbs need the same 100 32-bit multiplies per field multiplication and none of the glue.
// All arithmetic is exact; parity with the reference is defined on canonical encodings
// (field.rs:368-450 `to_bytes`), which fe_tobytes() reproduces bit for bit.
//
// Limb bounds ("scale" s means even limbs <= s*2^26, odd limbs <= s*2^25, with ~1% slack):
//   fe_mul / fe_sq outputs            : scale 1
//   fe_add(a,b)                        : scale(a)+scale(b), no carry
//   fe_sub(a,b)   (b scale <= 1)       : scale(a)+2      (adds 2p)
//   fe_sub2(a,b)  (b scale <= 2)       : scale(a)+4      (adds 4p)
//   fe_mul(f,g) requires scale(g) <= 3.3 (19*g_j must fit 32 bits) and scale(f)*scale(g) <= 30
//   fe_sq(f)   requires scale(f) <= 2
// The host build of this header (tests/) checks these bounds on every call when FE_CHECK_BOUNDS
// is defined.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#define FE_DEVCONST static __device__ __constant__
#else
#define FE_HD inline
#define FE_DEVCONST static const
#ifndef __align__
#define __align__(n) alignas(n)
#endif
#endif

#if defined(FE_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
#include <assert.h>
#define FE_ASSERT_SCALE(f, s) fe_assert_scale((f), (s))
#else
#define FE_ASSERT_SCALE(f, s) ((void)0)
#endif

struct fe { uint32_t v[10]; };

#if defined(FE_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
static inline void fe_assert_scale(const fe &f, double s)
{
    for (int i = 0; i < 10; i++) {
        double lim = s * 1.01 * (double)(1u << ((i & 1) ? 25 : 26));
        assert((double)f.v[i] <= lim);
    }
}
#endif

#define FE_M26 0x3ffffffu
#define FE_M25 0x1ffffffu

FE_HD void fe_0(fe &h) { for (int i = 0; i < 10; i++) h.v[i] = 0; }
FE_HD void fe_1(fe &h) { h.v[0] = 1; for (int i = 1; i < 10; i++) h.v[i] = 0; }
FE_HD void fe_copy(fe &h, const fe &f) { for (int i = 0; i < 10; i++) h.v[i] = f.v[i]; }

FE_HD void fe_add(fe &h, const fe &f, const fe &g)
{
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + g.v[i];
}

// h = f - g + 2p; g must have scale <= 1
FE_HD void fe_sub(fe &h, const fe &f, const fe &g)
{
    FE_ASSERT_SCALE(g, 1.0);
    h.v[0] = f.v[0] + 0x7ffffdau - g.v[0];
#pragma unroll
    for (int i = 1; i < 10; i++) h.v[i] = f.v[i] + ((i & 1) ? 0x3fffffeu : 0x7fffffeu) - g.v[i];
}

// h = f - g + 4p; g must have scale <= 2
FE_HD void fe_sub2(fe &h, const fe &f, const fe &g)
{
    FE_ASSERT_SCALE(g, 2.0);
    h.v[0] = f.v[0] + 0xfffffb4u - g.v[0];
#pragma unroll
    for (int i = 1; i < 10; i++) h.v[i] = f.v[i] + ((i & 1) ? 0x7fffffcu : 0xffffffcu) - g.v[i];
}

// h = -f (+2p); f scale <= 1
FE_HD void fe_neg(fe &h, const fe &f)
{
    FE_ASSERT_SCALE(f, 1.0);
    h.v[0] = 0x7ffffdau - f.v[0];
#pragma unroll
    for (int i = 1; i < 10; i++) h.v[i] = ((i & 1) ? 0x3fffffeu : 0x7fffffeu) - f.v[i];
}

// Carry ten 64-bit columns down to scale 1 (two interleaved chains, as in the classic ref10
// schedule, for instruction-level parallelism).
FE_HD void fe_carry64(fe &h, uint64_t c[10])
{
    uint64_t t;
    t = c[0] >> 26; c[1] += t; c[0] &= FE_M26;
    t = c[4] >> 26; c[5] += t; c[4] &= FE_M26;
    t = c[1] >> 25; c[2] += t; c[1] &= FE_M25;
    t = c[5] >> 25; c[6] += t; c[5] &= FE_M25;
    t = c[2] >> 26; c[3] += t; c[2] &= FE_M26;
    t = c[6] >> 26; c[7] += t; c[6] &= FE_M26;
    t = c[3] >> 25; c[4] += t; c[3] &= FE_M25;
    t = c[7] >> 25; c[8] += t; c[7] &= FE_M25;
    t = c[4] >> 26; c[5] += t; c[4] &= FE_M26;
    t = c[8] >> 26; c[9] += t; c[8] &= FE_M26;
    t = c[9] >> 25; c[0] += t * 19; c[9] &= FE_M25;
    t = c[0] >> 26; c[1] += t; c[0] &= FE_M26;
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = (uint32_t)c[i];
}

// Weak carry of 32-bit limbs back to scale 1 (input any scale that fits 32 bits).
FE_HD void fe_carry(fe &h, const fe &f)
{
    uint64_t c[10];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = f.v[i];
    fe_carry64(h, c);
}

#define FE_MAC(acc, a, b) (acc) += (uint64_t)(a) * (uint64_t)(b)

// h = f * g.  100 IMAD.WIDE.U32 + 9 multiplies by 19 + 5 doublings + the carry chain.
FE_HD void fe_mul(fe &h, const fe &f, const fe &g)
{
    FE_ASSERT_SCALE(g, 3.3);
    uint32_t g19[10], f2[10];
#pragma unroll
    for (int j = 1; j < 10; j++) g19[j] = g.v[j] * 19u;
#pragma unroll
    for (int i = 1; i < 10; i += 2) f2[i] = f.v[i] * 2u;
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const int k = (i + j) % 10;
            const bool wrap = (i + j) >= 10;
            const bool both_odd = (i & 1) && (j & 1);
            const uint32_t a = both_odd ? f2[i] : f.v[i];
            const uint32_t b = wrap ? g19[j] : g.v[j];
            FE_MAC(c[k], a, b);
        }
    }
    fe_carry64(h, c);
}

// h = f^2.  55 IMAD.WIDE.U32.
FE_HD void fe_sq(fe &h, const fe &f)
{
    FE_ASSERT_SCALE(f, 2.0);
    uint32_t f2[10], f19[10], f38[10];
#pragma unroll
    for (int i = 0; i < 10; i++) f2[i] = f.v[i] * 2u;
#pragma unroll
    for (int j = 5; j < 10; j++) { f19[j] = f.v[j] * 19u; f38[j] = f.v[j] * 38u; }
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
#pragma unroll
        for (int j = i; j < 10; j++) {
            const int k = (i + j) % 10;
            const bool wrap = (i + j) >= 10;
            const bool both_odd = (i & 1) && (j & 1);
            const uint32_t a = (i < j) ? f2[i] : f.v[i];
            // right factor: x1, x2 (both odd), x19 (wrap), x38 (wrap and both odd)
            const uint32_t b = wrap ? (both_odd ? f38[j] : f19[j]) : (both_odd ? f2[j] : f.v[j]);
            FE_MAC(c[k], a, b);
        }
    }
    fe_carry64(h, c);
}

// h = 2 * f^2 (reference `square2`, field.rs:567-574)
FE_HD void fe_sq2(fe &h, const fe &f)
{
    fe t; fe_sq(t, f);
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = t.v[i] * 2u;
}

// h = f * small constant (< 2^31 / scale)
FE_HD void fe_mul_small(fe &h, const fe &f, uint32_t k)
{
    uint64_t c[10];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = (uint64_t)f.v[i] * k;
    fe_carry64(h, c);
}

FE_HD void fe_sqn(fe &h, const fe &f, int n)
{
    fe_sq(h, f);
    for (int i = 1; i < n; i++) fe_sq(h, h);
}

// Unpack 32 little-endian bytes given as eight 32-bit words; bit 255 is ignored
// (field.rs:338-363 `from_bytes` semantics: no canonicity check).
FE_HD void fe_frombytes_words(fe &h, const uint32_t w[8])
{
    // limb i starts at bit offset ceil(25.5*i): 0,26,51,77,102,128,153,179,204,230
    h.v[0] = w[0] & FE_M26;
    h.v[1] = ((w[0] >> 26) | (w[1] << 6)) & FE_M25;
    h.v[2] = ((w[1] >> 19) | (w[2] << 13)) & FE_M26;
    h.v[3] = ((w[2] >> 13) | (w[3] << 19)) & FE_M25;
    h.v[4] = (w[3] >> 6) & FE_M26;
    h.v[5] = w[4] & FE_M25;
    h.v[6] = ((w[4] >> 25) | (w[5] << 7)) & FE_M26;
    h.v[7] = ((w[5] >> 19) | (w[6] << 13)) & FE_M25;
    h.v[8] = ((w[6] >> 12) | (w[7] << 20)) & FE_M26;
    h.v[9] = (w[7] >> 6) & FE_M25;
}

// Canonical encoding into eight 32-bit words (bit 255 clear); field.rs:368-450.
FE_HD void fe_tobytes_words(uint32_t w[8], const fe &f)
{
    fe t; fe_carry(t, f);          // scale 1, value < 2p + small
    uint32_t *v = t.v;
    // q = (value + 19) >> 255, computed by propagating the carry of +19 through the limbs
    uint32_t q = (v[0] + 19) >> 26;
#pragma unroll
    for (int i = 1; i < 10; i++) q = (v[i] + q) >> ((i & 1) ? 25 : 26);
    v[0] += 19 * q;
    uint32_t c;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c = v[i] >> ((i & 1) ? 25 : 26); v[i + 1] += c; v[i] &= (i & 1) ? FE_M25 : FE_M26;
    }
    v[9] &= FE_M25;
    // the value may still have been >= 2p - ... only if the input was not weakly reduced; one
    // weak carry above guarantees value < 2^255 + 2^230ish < 2p, so a single conditional
    // subtraction (the +19 trick) is exact.
    w[0] = v[0] | (v[1] << 26);
    w[1] = (v[1] >> 6) | (v[2] << 19);
    w[2] = (v[2] >> 13) | (v[3] << 13);
    w[3] = (v[3] >> 19) | (v[4] << 6);
    w[4] = v[5] | (v[6] << 25);
    w[5] = (v[6] >> 7) | (v[7] << 19);
    w[6] = (v[7] >> 13) | (v[8] << 12);
    w[7] = (v[8] >> 20) | (v[9] << 6);
}

// 1 if the canonical encodings are equal (C/field.rs:92-99)
FE_HD int fe_eq(const fe &a, const fe &b)
{
    uint32_t x[8], y[8];
    fe_tobytes_words(x, a); fe_tobytes_words(y, b);
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= x[i] ^ y[i];
    return d == 0;
}

FE_HD int fe_isnegative(const fe &a) { uint32_t x[8]; fe_tobytes_words(x, a); return x[0] & 1; }

FE_HD int fe_iszero(const fe &a)
{
    uint32_t x[8]; fe_tobytes_words(x, a);
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= x[i];
    return d == 0;
}

// branch-free select: h = c ? g : h   (c in {0,1})
FE_HD void fe_cmov(fe &h, const fe &g, uint32_t c)
{
    uint32_t m = 0u - c;
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] ^= m & (h.v[i] ^ g.v[i]);
}

// h = c ? -h : h, input scale <= 1, output scale <= 2 (not carried)
FE_HD void fe_cneg(fe &h, uint32_t c)
{
    fe n; fe_neg(n, h); fe_cmov(h, n, c);
}

// (f^(2^250-1), f^11): C/field.rs:176-210 addition chain
FE_HD void fe_pow22501(fe &t19, fe &t3, const fe &f)
{
    fe t0, t1, t2, t4, t5, t6, t7, t9, t13, t15;
    fe_sq(t0, f);
    fe_sqn(t1, t0, 2);
    fe_mul(t2, f, t1);
    fe_mul(t3, t0, t2);
    fe_sq(t4, t3);
    fe_mul(t5, t2, t4);
    fe_sqn(t6, t5, 5);   fe_mul(t7, t6, t5);
    fe_sqn(t6, t7, 10);  fe_mul(t9, t6, t7);
    fe_sqn(t6, t9, 20);  fe_mul(t6, t6, t9);
    fe_sqn(t6, t6, 10);  fe_mul(t13, t6, t7);
    fe_sqn(t6, t13, 50); fe_mul(t15, t6, t13);
    fe_sqn(t6, t15, 100); fe_mul(t6, t6, t15);
    fe_sqn(t6, t6, 50);  fe_mul(t19, t6, t13);
}

// f^(p-2): C/field.rs:283-292
FE_HD void fe_invert(fe &h, const fe &f)
{
    fe t19, t3;
    fe_pow22501(t19, t3, f);
    fe_sqn(t19, t19, 5);
    fe_mul(h, t19, t3);
}

// f^((p-5)/8): C/field.rs:297-306
FE_HD void fe_pow_p58(fe &h, const fe &f)
{
    fe t19, t3;
    fe_pow22501(t19, t3, f);
    fe_sqn(t19, t19, 2);
    fe_mul(h, f, t19);
}
