// fe64.cuh -- GF(2^255-19) on the FP64 pipe of sm_100a, used by the throughput-bound bucket kernel.
//
// Measured on B200 (profiles/microbench_f64_r1.json): DFMA issues at 64 /clk/SM, IMAD.WIDE.U32 at
// 32 /clk/SM; a field multiplication built on DFMA runs at 112-119 G/s against 70 G/s for the
// IMAD.WIDE form in fe.cuh.  Elements are five integer-valued doubles in radix 2^51, kept BALANCED
// (|limb| <= 2^50 after a multiplication) so that sums of two or three elements are still valid
// multiplication operands without a carry.
//
// A limb product p = a_i * b_j (|p| < 2^103) is split exactly into p = f * 2^52 + lo with
//     t  = fma_rz(a_i, b_j * 2^-52, 1.5 * 2^52)   = 1.5 * 2^52 + f,        f = floor(p / 2^52)
//     u  = fma_rn(t, -2^52, 1.5 * 2^104 + 2^52)   = 2^52 - f * 2^52        (exact)
//     lo'= fma_rn(a_i, b_j, u)                    = 2^52 + (p mod 2^52)    (exact)
// (round-toward-zero on a positive sum is floor; both t and lo' lie in [2^52, 2^53), where a double's
// mantissa field IS the integer offset).  The raw IEEE bit patterns of t and lo' are accumulated per
// column with 64-bit integer adds; the known exponent offsets are removed once per column.  The nine
// columns are then wrapped (2^255 = 19), carried to balanced limbs in integer arithmetic and turned
// back into doubles.  Everything is exact; results are converted to the integer representation
// (fe.cuh) before they leave the kernel, so canonical encodings are unchanged.
//
// Operand rule: for every limb pair |a_i| * |b_j| < 2^103.  "Scale" s below means |limb| <= s * 2^50
// (+ s * 2^15 slack: fe64_finish leaves 2^10, the one-round carry of the 20-lane code in warp4_f64.cuh 2^14): fe64_mul
// needs scale(a) * scale(b) < 8; outputs have scale 1.
#pragma once
#include <stdint.h>

#include "fe.cuh"

struct fe64 { double v[5]; };

#define FE64_E52 0x4330000000000000LL          // bit pattern of 2^52
#define FE64_TWO52 4503599627370496.0
#define FE64_TWO51 2251799813685248.0

#if defined(__CUDA_ARCH__)
#define FE64_DEV 1
#else
#define FE64_DEV 0
#include <assert.h>
#include <math.h>
#include <string.h>
#endif

FE_HD void fe64_0(fe64 &h) { for (int i = 0; i < 5; i++) h.v[i] = 0.0; }
FE_HD void fe64_1(fe64 &h) { h.v[0] = 1.0; for (int i = 1; i < 5; i++) h.v[i] = 0.0; }

FE_HD void fe64_add(fe64 &h, const fe64 &f, const fe64 &g)
{
#pragma unroll
    for (int i = 0; i < 5; i++) h.v[i] = f.v[i] + g.v[i];
}
FE_HD void fe64_sub(fe64 &h, const fe64 &f, const fe64 &g)
{
#pragma unroll
    for (int i = 0; i < 5; i++) h.v[i] = f.v[i] - g.v[i];
}

// h = c ? g : h   (c in {0,1}), branch-free
FE_HD void fe64_cmov(fe64 &h, const fe64 &g, uint32_t c)
{
#if FE64_DEV
#pragma unroll
    for (int i = 0; i < 5; i++) {
        long long a = __double_as_longlong(h.v[i]), b = __double_as_longlong(g.v[i]);
        long long m = 0LL - (long long)c;
        h.v[i] = __longlong_as_double(a ^ (m & (a ^ b)));
    }
#else
    if (c) h = g;
#endif
}

#if !FE64_DEV
static inline void fe64_assert_scale(const fe64 &f, double s)
{
    for (int i = 0; i < 5; i++) {
        assert(f.v[i] == floor(f.v[i]));
        assert(fabs(f.v[i]) <= s * 1125899906842624.0 + 32768.0 * s);
    }
}
#define FE64_ASSERT_SCALE(f, s) fe64_assert_scale((f), (s))
#else
#define FE64_ASSERT_SCALE(f, s) ((void)0)
#endif

// nine signed 64-bit columns (weight 2^(51k)) -> wrap, balanced carry, doubles
FE_HD void fe64_finish(fe64 &h, long long V[9])
{
    long long R[5];
#pragma unroll
    for (int k = 0; k < 4; k++) R[k] = V[k] + 19 * V[k + 5];
    R[4] = V[4];
    const long long HALF = 1LL << 50;
    long long c;
    c = (R[0] + HALF) >> 51; R[0] -= c * (1LL << 51); R[1] += c;
    c = (R[1] + HALF) >> 51; R[1] -= c * (1LL << 51); R[2] += c;
    c = (R[2] + HALF) >> 51; R[2] -= c * (1LL << 51); R[3] += c;
    c = (R[3] + HALF) >> 51; R[3] -= c * (1LL << 51); R[4] += c;
    c = (R[4] + HALF) >> 51; R[4] -= c * (1LL << 51); R[0] += 19 * c;
    c = (R[0] + HALF) >> 51; R[0] -= c * (1LL << 51); R[1] += c;
#pragma unroll
    for (int k = 0; k < 5; k++) {
#if FE64_DEV
        h.v[k] = __longlong_as_double((R[k] + (1LL << 51)) | FE64_E52) - (FE64_TWO52 + FE64_TWO51);
#else
        h.v[k] = (double)R[k];
#endif
    }
}

// h = a * b.   25 x (2 DFMA.RZ/RN + 1 DFMA) + 5 DMUL on the FP64 pipe, 50 64-bit integer adds.
FE_HD void fe64_mul(fe64 &h, const fe64 &a, const fe64 &b)
{
    long long V[9];
#if FE64_DEV
    const double M1 = 6755399441055744.0;                               // 1.5 * 2^52
    const double K = 6755399441055744.0 * 4503599627370496.0 + 4503599627370496.0;   // M1 * 2^52 + 2^52
    double bs[5];
#pragma unroll
    for (int j = 0; j < 5; j++) bs[j] = b.v[j] * (1.0 / 4503599627370496.0);
    long long H[10], L[9];
#pragma unroll
    for (int k = 0; k < 10; k++) H[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) L[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double t = __fma_rz(a.v[i], bs[j], M1);
            double u = __fma_rn(t, -4503599627370496.0, K);
            double lo = __fma_rn(a.v[i], b.v[j], u);
            H[i + j + 1] += __double_as_longlong(t);
            L[i + j] += __double_as_longlong(lo);
        }
    const long long EH = FE64_E52 + (1LL << 51);                         // bits(1.5 * 2^52)
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const int nl = k < 5 ? k + 1 : 9 - k;
        const int nh = k == 0 ? 0 : (k <= 5 ? k : 10 - k);
        V[k] = (L[k] - nl * FE64_E52) + 2 * (H[k] - nh * EH);           // lo + 2 f (f has weight 2^52 = 2 * 2^51)
    }
    V[4] += 38 * (H[9] - EH);                                            // column 9 (only f of a4*b4): 2 * 19
#else
    // host model of the same exact arithmetic (tests only): checks the operand rule
    for (int k = 0; k < 9; k++) V[k] = 0;
    long long V9 = 0;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) {
            __int128 p = (__int128)(long long)a.v[i] * (__int128)(long long)b.v[j];
            __int128 lim = (__int128)1 << 103;
            assert(p < lim && p > -lim);
            long long f = (long long)(p >> 52);                          // floor
            long long lo = (long long)(p - ((__int128)f << 52));
            V[i + j] += lo;
            if (i + j + 1 < 9) V[i + j + 1] += 2 * f; else V9 += 2 * f;
        }
    V[4] += 19 * V9;
#endif
    fe64_finish(h, V);
}

// h = a^2.   15 x (2 DFMA.RZ/RN + 1 DFMA): cross terms go through a pre-doubled operand, so the operand
// rule is 2 |a_i a_j| < 2^103, i.e. scale(a) < 2 (every product or square qualifies).
FE_HD void fe64_sq(fe64 &h, const fe64 &a)
{
    long long V[9];
#if FE64_DEV
    const double M1 = 6755399441055744.0;                               // 1.5 * 2^52
    const double K = 6755399441055744.0 * 4503599627370496.0 + 4503599627370496.0;
    double as[5], a2[5];
#pragma unroll
    for (int j = 0; j < 5; j++) { as[j] = a.v[j] * (1.0 / 4503599627370496.0); a2[j] = a.v[j] + a.v[j]; }
    long long H[10], L[9];
#pragma unroll
    for (int k = 0; k < 10; k++) H[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) L[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = i; j < 5; j++) {
            const double x = i < j ? a2[i] : a.v[i];
            double t = __fma_rz(x, as[j], M1);
            double u = __fma_rn(t, -4503599627370496.0, K);
            double lo = __fma_rn(x, a.v[j], u);
            H[i + j + 1] += __double_as_longlong(t);
            L[i + j] += __double_as_longlong(lo);
        }
    const long long EH = FE64_E52 + (1LL << 51);
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const int nl = k < 5 ? k / 2 + 1 : (8 - k) / 2 + 1;            // pairs i <= j with i + j = k
        const int km = k - 1;
        const int nh = k == 0 ? 0 : (km < 5 ? km / 2 + 1 : (8 - km) / 2 + 1);
        V[k] = (L[k] - nl * FE64_E52) + 2 * (H[k] - nh * EH);
    }
    V[4] += 38 * (H[9] - EH);
#else
    for (int k = 0; k < 9; k++) V[k] = 0;
    long long V9 = 0;
    for (int i = 0; i < 5; i++)
        for (int j = i; j < 5; j++) {
            __int128 p = (__int128)(long long)a.v[i] * (__int128)(long long)a.v[j] * (i < j ? 2 : 1);
            __int128 lim = (__int128)1 << 103;
            assert(p < lim && p > -lim);
            long long f = (long long)(p >> 52);
            long long lo = (long long)(p - ((__int128)f << 52));
            V[i + j] += lo;
            if (i + j + 1 < 9) V[i + j + 1] += 2 * f; else V9 += 2 * f;
        }
    V[4] += 19 * V9;
#endif
    fe64_finish(h, V);
}

FE_HD void fe64_sqn(fe64 &h, const fe64 &f, int n)
{
    fe64_sq(h, f);
#if FE64_DEV
#pragma unroll 1
#endif
    for (int i = 1; i < n; i++) fe64_sq(h, h);
}

// weak balanced carry in the FP domain: output scale 1 (+ a few units), input any scale <= 2^12
FE_HD void fe64_carry(fe64 &h, const fe64 &f)
{
#if FE64_DEV
    const double C = 6755399441055744.0;                                 // 1.5 * 2^52
    double q[5], r[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        q[k] = __fma_rn(f.v[k], 1.0 / 2251799813685248.0, C) - C;       // round(f / 2^51)
        r[k] = __fma_rn(q[k], -2251799813685248.0, f.v[k]);
    }
    h.v[0] = __fma_rn(q[4], 19.0, r[0]);
#pragma unroll
    for (int k = 1; k < 5; k++) h.v[k] = r[k] + q[k - 1];
#else
    double q[5], r[5];
    for (int k = 0; k < 5; k++) { q[k] = nearbyint(f.v[k] / 2251799813685248.0); r[k] = f.v[k] - q[k] * 2251799813685248.0; }
    h.v[0] = r[0] + 19.0 * q[4];
    for (int k = 1; k < 5; k++) h.v[k] = r[k] + q[k - 1];
#endif
}

// 32 canonical little-endian bytes as eight 32-bit words -> limbs in [0, 2^51)  (scale 2)
FE_HD void fe64_frombytes_words(fe64 &h, const uint32_t w[8])
{
    uint64_t l[5];
    l[0] = (uint64_t)w[0] | ((uint64_t)(w[1] & 0x7ffffu) << 32);
    l[1] = (uint64_t)(w[1] >> 19) | ((uint64_t)w[2] << 13) | ((uint64_t)(w[3] & 0x3fu) << 45);
    l[2] = (uint64_t)(w[3] >> 6) | ((uint64_t)(w[4] & 0x1ffffffu) << 26);
    l[3] = (uint64_t)(w[4] >> 25) | ((uint64_t)w[5] << 7) | ((uint64_t)(w[6] & 0xfffu) << 39);
    l[4] = (uint64_t)(w[6] >> 12) | ((uint64_t)(w[7] & 0x7fffffffu) << 20);
#pragma unroll
    for (int k = 0; k < 5; k++) {
#if FE64_DEV
        h.v[k] = __longlong_as_double((long long)l[k] | FE64_E52) - FE64_TWO52;
#else
        h.v[k] = (double)l[k];
#endif
    }
}

// to the integer representation of fe.cuh (any scale <= 4)
FE_HD void fe64_to_fe(fe &o, const fe64 &f)
{
    // add 8p limb-wise so that every limb is positive, then split 51 -> 26 + 25 bits
    uint64_t l[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        long long r = (long long)f.v[k];
        l[k] = (uint64_t)(r + (k == 0 ? (1LL << 54) - 152 : (1LL << 54) - 8));
    }
    uint64_t c[10];
#pragma unroll
    for (int i = 0; i < 5; i++) { c[2 * i] = l[i] & FE_M26; c[2 * i + 1] = l[i] >> 26; }
    fe_carry64(o, c);
}

FE_HD void fe64_from_fe(fe64 &h, const fe &f)
{
    uint32_t w[8];
    fe_tobytes_words(w, f);
    fe64 t; fe64_frombytes_words(t, w);
    fe64_carry(h, t);
}

// (f^(2^250-1), f^11) with the addition chain of C/field.rs:176-210 on the FP64 field
FE_HD void fe64_pow22501(fe64 &t19, fe64 &t3, const fe64 &z)
{
    fe64 t0, t1, t2, t5, t6, t7, t9, t13, t15;
    fe64_sq(t0, z);
    fe64_sqn(t1, t0, 2);
    fe64_mul(t2, z, t1);
    fe64_mul(t3, t0, t2);
    fe64_sq(t1, t3);
    fe64_mul(t5, t2, t1);
    fe64_sqn(t6, t5, 5);    fe64_mul(t7, t6, t5);
    fe64_sqn(t6, t7, 10);   fe64_mul(t9, t6, t7);
    fe64_sqn(t6, t9, 20);   fe64_mul(t6, t6, t9);
    fe64_sqn(t6, t6, 10);   fe64_mul(t13, t6, t7);
    fe64_sqn(t6, t13, 50);  fe64_mul(t15, t6, t13);
    fe64_sqn(t6, t15, 100); fe64_mul(t6, t6, t15);
    fe64_sqn(t6, t6, 50);   fe64_mul(t19, t6, t13);
}

// f^((p-5)/8) (C/field.rs:297-306) on the FP64 field: 252 squarings at the FP64 squaring rate (142 G/s against
// 126 G/s for the IMAD.WIDE form on B200).
FE_HD void fe_pow_p58_f64(fe &h, const fe &f)
{
    fe64 z, t19, t3;
    fe64_from_fe(z, f);
    fe64_pow22501(t19, t3, z);
    fe64_sqn(t19, t19, 2);
    fe64_mul(t19, z, t19);
    fe64_to_fe(h, t19);
}

// f^(p-2) (C/field.rs:283-292) on the FP64 field; 0 -> 0
FE_HD void fe_invert_f64(fe &h, const fe &f)
{
    fe64 z, t19, t3;
    fe64_from_fe(z, f);
    fe64_pow22501(t19, t3, z);
    fe64_sqn(t19, t19, 5);
    fe64_mul(t19, t19, t3);
    fe64_to_fe(h, t19);
}
