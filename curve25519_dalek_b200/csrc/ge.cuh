// ge.cuh -- twisted Edwards point arithmetic (a = -1) over fe.cuh for sm_100a.
//
// Same curve models and formulas as the reference
// (curve25519-dalek/src/backend/serial/curve_models.rs:154-523, src/edwards.rs:211-257,
// :528-617, :786-871, :1370-1380) but with the conversions fused (completed -> extended is
// folded into the add) and with limb-scale bookkeeping for the 32-bit limbs (see fe.cuh).
// The unified addition is complete on this curve, so identity / doubling / torsion inputs
// need no special cases.
#pragma once
#include "constants.cuh"
#include "fe.cuh"
#include "fe64.cuh"

struct ge_p3 { fe X, Y, Z, T; };           // EdwardsPoint (extended), all scale 1
struct ge_p2 { fe X, Y, Z; };              // ProjectivePoint
struct ge_niels { fe ypx, ymx, xy2d; };    // AffineNielsPoint (y+x, y-x, 2dxy), scale <= 1
struct ge_pniels { fe YpX, YmX, Z, T2d; }; // ProjectiveNielsPoint, scale <= 1 (carried)

// 32-byte-per-coordinate storage formats in HBM (canonical little-endian field encodings)
struct __align__(16) ge_niels_packed { uint32_t w[24]; };   //  96 B: ypx | ymx | xy2d
struct __align__(16) ge_pniels_packed { uint32_t w[32]; };  // 128 B: YpX | YmX | Z | T2d
struct __align__(16) ge_p3_raw { uint32_t w[40]; };         // 160 B: X | Y | Z | T as 10 limbs each

FE_HD void ge_p3_identity(ge_p3 &p) { fe_0(p.X); fe_1(p.Y); fe_1(p.Z); fe_0(p.T); }

FE_HD void ge_p3_basepoint(ge_p3 &p)
{
    fe_const_base_x(p.X); fe_const_base_y(p.Y); fe_1(p.Z); fe_const_base_t(p.T);
}

// r = p + q (q affine Niels), or p - q when neg = 1.   7M.
// curve_models.rs:455-494 followed by :365-372.  Branch-free in `neg`.
FE_HD void ge_madd(ge_p3 &r, const ge_p3 &p, const ge_niels &q, uint32_t neg)
{
    fe A, B, a, b, c, D, E, H, DpC, DmC, F, G;
    fe qp = q.ypx, qm = q.ymx;
    {   // -q swaps (y+x) and (y-x) and negates xy2d (curve_models.rs:514-523)
        fe t = qp; fe_cmov(qp, qm, neg); fe_cmov(qm, t, neg);
    }
    fe_sub(A, p.Y, p.X);            // 3
    fe_add(B, p.Y, p.X);            // 2
    fe_mul(a, A, qm);               // 1
    fe_mul(b, B, qp);               // 1
    fe_mul(c, p.T, q.xy2d);         // 1
    fe_add(D, p.Z, p.Z);            // 2
    fe_sub(E, b, a);                // 3   X of completed
    fe_add(H, b, a);                // 2   Y of completed
    fe_add(DpC, D, c);              // 3
    fe_sub(DmC, D, c);              // 4
    F = DmC; fe_cmov(F, DpC, neg);  // T of completed (Z - c, or Z + c when subtracting)
    G = DpC; fe_cmov(G, DmC, neg);  // Z of completed
    fe_mul(r.X, F, E);              // f scale <= 4, g scale 3
    fe_mul(r.Y, G, H);              // <= 4, 2
    fe_mul(r.Z, DmC, DpC);          // 4, 3  (F*G in either order)
    fe_mul(r.T, E, H);              // 3, 2
}

// r = p + q (q projective Niels), or p - q when neg = 1.   8M.
// curve_models.rs:411-452 followed by :365-372.
FE_HD void ge_padd(ge_p3 &r, const ge_p3 &p, const ge_pniels &q, uint32_t neg)
{
    fe A, B, a, b, c, ZZ, D, E, H, DpC, DmC, F, G;
    fe qp = q.YpX, qm = q.YmX;
    {
        fe t = qp; fe_cmov(qp, qm, neg); fe_cmov(qm, t, neg);
    }
    fe_sub(A, p.Y, p.X);
    fe_add(B, p.Y, p.X);
    fe_mul(a, A, qm);
    fe_mul(b, B, qp);
    fe_mul(c, p.T, q.T2d);
    fe_mul(ZZ, p.Z, q.Z);
    fe_add(D, ZZ, ZZ);
    fe_sub(E, b, a);
    fe_add(H, b, a);
    fe_add(DpC, D, c);
    fe_sub(DmC, D, c);
    F = DmC; fe_cmov(F, DpC, neg);
    G = DpC; fe_cmov(G, DmC, neg);
    fe_mul(r.X, F, E);
    fe_mul(r.Y, G, H);
    fe_mul(r.Z, DmC, DpC);
    fe_mul(r.T, E, H);
}

// Completed point of a doubling (curve_models.rs:381-397): X' (carried to scale 1), Y' (2),
// Z' (3), T' (5).   4S + one weak carry.
struct ge_p1p1 { fe X, Y, Z, T; };
FE_HD void ge_dbl_p1p1(ge_p1p1 &r, const fe &X, const fe &Y, const fe &Z)
{
    fe XX, YY, ZZ2, S, t;
    fe_sq(XX, X);
    fe_sq(YY, Y);
    fe_sq2(ZZ2, Z);                 // 2
    fe_add(S, X, Y);                // 2
    fe_sq(S, S);                    // 1   (X+Y)^2
    fe_sub(t, S, YY);               // 3
    fe_sub(t, t, XX);               // 5
    fe_carry(r.X, t);               // 1   X' = (X+Y)^2 - YY - XX
    fe_add(r.Y, YY, XX);            // 2   Y' = YY + XX
    fe_sub(r.Z, YY, XX);            // 3   Z' = YY - XX
    fe_add(t, ZZ2, XX);             // 3
    fe_sub(r.T, t, YY);             // 5   T' = 2ZZ - (YY - XX)
}

// completed -> extended (curve_models.rs:365-372), 4M
FE_HD void ge_p1p1_to_p3(ge_p3 &r, const ge_p1p1 &c)
{
    fe_mul(r.X, c.T, c.X);          // f=T'(5), g=X'(1)
    fe_mul(r.Y, c.Z, c.Y);          // 3, 2
    fe_mul(r.T, c.Y, c.X);          // 2, 1   (before Z: r may alias nothing here)
    fe_mul(r.Z, c.T, c.Z);          // 5, 3
}

// completed -> projective (curve_models.rs:353-359), 3M
FE_HD void ge_p1p1_to_p2(ge_p2 &r, const ge_p1p1 &c)
{
    fe_mul(r.X, c.T, c.X);
    fe_mul(r.Y, c.Z, c.Y);
    fe_mul(r.Z, c.T, c.Z);
}

// r = 2p   (C/edwards.rs:786-788), 4S + 4M
FE_HD void ge_dbl(ge_p3 &r, const ge_p3 &p)
{
    ge_p1p1 c; ge_dbl_p1p1(c, p.X, p.Y, p.Z); ge_p1p1_to_p3(r, c);
}

// r = 2^k p, k >= 1   (C/edwards.rs:1370-1380): k-1 projective doublings (4S+3M) and a final
// extended one (4S+4M)
FE_HD void ge_mul_by_pow_2(ge_p3 &r, const ge_p3 &p, int k)
{
    ge_p2 s; s.X = p.X; s.Y = p.Y; s.Z = p.Z;
    ge_p1p1 c;
    for (int i = 0; i + 1 < k; i++) { ge_dbl_p1p1(c, s.X, s.Y, s.Z); ge_p1p1_to_p2(s, c); }
    ge_dbl_p1p1(c, s.X, s.Y, s.Z);
    ge_p1p1_to_p3(r, c);
}

// EdwardsPoint::as_projective_niels (C/edwards.rs:528-535), carried so it can be stored / reused
FE_HD void ge_p3_to_pniels(ge_pniels &r, const ge_p3 &p)
{
    fe t, d2;
    fe_const_2d(d2);
    fe_add(t, p.Y, p.X); fe_carry(r.YpX, t);
    fe_sub(t, p.Y, p.X); fe_carry(r.YmX, t);
    r.Z = p.Z;
    fe_mul(r.T2d, p.T, d2);
}

// r = p + q, both extended (C/edwards.rs:795-800): 1M + 8M
FE_HD void ge_add(ge_p3 &r, const ge_p3 &p, const ge_p3 &q)
{
    ge_pniels n; ge_p3_to_pniels(n, q); ge_padd(r, p, n, 0);
}

// affine (x, y) -> AffineNielsPoint (curve_models.rs:184-189)
FE_HD void ge_affine_to_niels(ge_niels &r, const fe &x, const fe &y)
{
    fe t, d2;
    fe_const_2d(d2);
    fe_add(t, y, x); fe_carry(r.ypx, t);
    fe_sub(t, y, x); fe_carry(r.ymx, t);
    fe_mul(t, x, y);
    fe_mul(r.xy2d, t, d2);
}

// sqrt_ratio_i (C/field.rs:320-366): returns was_nonzero_square, r = nonnegative root
// F64 = 1 runs the 252-squaring exponentiation on the FP64-pipe field (fe64.cuh): same value, faster on B200.
template <int F64 = 0>
FE_HD uint32_t fe_sqrt_ratio_i(fe &r, const fe &u, const fe &v)
{
    fe v3, v7, t, check, i, neg_u, neg_u_i, r_prime;
    fe_const_sqrtm1(i);
    fe_sq(t, v); fe_mul(v3, t, v);
    fe_sq(t, v3); fe_mul(v7, t, v);
    fe_mul(t, u, v7);
    if (F64) fe_pow_p58_f64(t, t); else fe_pow_p58(t, t);
    fe_mul(r, u, v3); fe_mul(r, r, t);
    fe_sq(t, r); fe_mul(check, v, t);
    fe uc; fe_carry(uc, u);
    fe_neg(neg_u, uc);                       // scale 2
    fe_mul(neg_u_i, neg_u, i);
    uint32_t correct = fe_eq(check, uc);
    uint32_t flipped = fe_eq(check, neg_u);
    uint32_t flipped_i = fe_eq(check, neg_u_i);
    fe_mul(r_prime, r, i);
    fe_cmov(r, r_prime, flipped | flipped_i);
    uint32_t negv = fe_isnegative(r);
    fe_cneg(r, negv);                        // scale <= 2
    fe_carry(r, r);
    return correct | flipped;
}

// CompressedEdwardsY::decompress (C/edwards.rs:211-257).  s = 8 little-endian words.
// Returns 1 and affine (x, y) on success.  Accepts non-canonical y like the reference.
template <int F64 = 0>
FE_HD uint32_t ge_decompress_affine(fe &x, fe &y, const uint32_t s[8])
{
    fe one, YY, u, v, d;
    fe_const_d(d); fe_1(one);
    fe_frombytes_words(y, s);
    fe_sq(YY, y);
    fe_sub(u, YY, one);                      // 3
    fe_mul(v, YY, d); fe_add(v, v, one);     // ~1
    uint32_t ok = fe_sqrt_ratio_i<F64>(x, u, v);
    fe_cneg(x, s[7] >> 31);
    fe_carry(x, x);
    return ok;
}

// EdwardsPoint::compress (C/edwards.rs:564-617, edwards/affine.rs:71-75) -> 8 words
template <int F64 = 0>
FE_HD void ge_compress(uint32_t s[8], const ge_p3 &p)
{
    fe recip, x, y;
    if (F64) fe_invert_f64(recip, p.Z); else fe_invert(recip, p.Z);
    fe_mul(x, p.X, recip);
    fe_mul(y, p.Y, recip);
    fe_tobytes_words(s, y);
    s[7] ^= (uint32_t)fe_isnegative(x) << 31;
}

// is_identity (C/traits.rs:41-48 via ct_eq with (0,1,1,0), C/edwards.rs:501-512): X == 0 and Y == Z
FE_HD uint32_t ge_is_identity(const ge_p3 &p)
{
    return (uint32_t)(fe_iszero(p.X) & fe_eq(p.Y, p.Z));
}

// ---- packed storage helpers ----
FE_HD void ge_niels_pack(ge_niels_packed &o, const ge_niels &n)
{
    fe_tobytes_words(o.w, n.ypx); fe_tobytes_words(o.w + 8, n.ymx); fe_tobytes_words(o.w + 16, n.xy2d);
}
FE_HD void ge_niels_unpack(ge_niels &n, const ge_niels_packed &o)
{
    fe_frombytes_words(n.ypx, o.w); fe_frombytes_words(n.ymx, o.w + 8); fe_frombytes_words(n.xy2d, o.w + 16);
}
FE_HD void ge_pniels_pack(ge_pniels_packed &o, const ge_pniels &n)
{
    fe_tobytes_words(o.w, n.YpX); fe_tobytes_words(o.w + 8, n.YmX);
    fe_tobytes_words(o.w + 16, n.Z); fe_tobytes_words(o.w + 24, n.T2d);
}
FE_HD void ge_pniels_unpack(ge_pniels &n, const ge_pniels_packed &o)
{
    fe_frombytes_words(n.YpX, o.w); fe_frombytes_words(n.YmX, o.w + 8);
    fe_frombytes_words(n.Z, o.w + 16); fe_frombytes_words(n.T2d, o.w + 24);
}
FE_HD void ge_p3_store_raw(ge_p3_raw &o, const ge_p3 &p)
{
#pragma unroll
    for (int i = 0; i < 10; i++) { o.w[i] = p.X.v[i]; o.w[10 + i] = p.Y.v[i]; o.w[20 + i] = p.Z.v[i]; o.w[30 + i] = p.T.v[i]; }
}
FE_HD void ge_p3_load_raw(ge_p3 &p, const ge_p3_raw &o)
{
#pragma unroll
    for (int i = 0; i < 10; i++) { p.X.v[i] = o.w[i]; p.Y.v[i] = o.w[10 + i]; p.Z.v[i] = o.w[20 + i]; p.T.v[i] = o.w[30 + i]; }
}

// Reference in-memory EdwardsPoint: 4 x FieldElement51 = 20 u64 limbs in radix 2^51, limbs may
// be as large as 2^54 (u64/field.rs:27-43).  Convert one coordinate to radix 2^25.5.
FE_HD void fe_from_limbs51(fe &h, const uint64_t l[5])
{
    uint64_t c[10];
#pragma unroll
    for (int i = 0; i < 5; i++) { c[2 * i] = l[i] & FE_M26; c[2 * i + 1] = l[i] >> 26; }
    fe_carry64(h, c);
}

// Canonical radix-2^51 limbs of a coordinate (what FieldElement51::from_bytes(to_bytes(x)) gives)
FE_HD void fe_to_limbs51(uint64_t l[5], const fe &f)
{
    uint32_t w[8];
    fe_tobytes_words(w, f);
    fe t; fe_frombytes_words(t, w);
#pragma unroll
    for (int i = 0; i < 5; i++) l[i] = (uint64_t)t.v[2 * i] | ((uint64_t)t.v[2 * i + 1] << 26);
}

// ------------------------------------------------------------------------------------------
// Ristretto255 (curve25519-dalek/src/ristretto.rs)

// CompressedRistretto::decompress (ristretto.rs:266-345).  Returns 1 and an extended point with
// Z = 1 on success.
template <int F64 = 0>
FE_HD uint32_t ristretto_decompress(ge_p3 &p, const uint32_t in[8])
{
    fe s, one, ss, u1, u2, u2_sqr, v, t, I, Dx, Dy, x, y, nd, d;
    fe_frombytes_words(s, in);
    uint32_t chk[8], diff = 0;
    fe_tobytes_words(chk, s);
#pragma unroll
    for (int i = 0; i < 8; i++) diff |= chk[i] ^ in[i];
    uint32_t canonical = diff == 0;                  // step_1: s < p and bit 255 clear
    uint32_t s_neg = chk[0] & 1;
    fe_1(one); fe_const_d(d);
    fe_sq(ss, s);
    fe_sub(u1, one, ss); fe_carry(u1, u1);           // 1 + a s^2  (a = -1)
    fe_add(u2, one, ss);                             // 1 - a s^2
    fe_sq(u2_sqr, u2);
    fe_neg(nd, d);
    fe_sq(t, u1); fe_mul(t, nd, t);
    fe_sub(v, t, u2_sqr);                            // a d u1^2 - u2^2
    fe_mul(t, v, u2_sqr);
    uint32_t ok = fe_sqrt_ratio_i<F64>(I, one, t);   // invsqrt (C/field.rs:380-382)
    fe_mul(Dx, I, u2);
    fe_mul(t, v, Dx); fe_mul(Dy, I, t);
    fe_add(t, s, s); fe_mul(x, t, Dx);
    fe_cneg(x, (uint32_t)fe_isnegative(x)); fe_carry(x, x);
    fe_mul(y, u1, Dy);
    fe_mul(t, x, y);
    uint32_t bad = (1u - canonical) | s_neg | (1u - ok) | (uint32_t)fe_isnegative(t) | (uint32_t)fe_iszero(y);
    p.X = x; p.Y = y; p.Z = one; p.T = t;
    return 1u - (bad & 1u);
}

// RistrettoPoint::compress (ristretto.rs:500-533)
template <int F64 = 0>
FE_HD void ristretto_compress(uint32_t out[8], const ge_p3 &p)
{
    fe X = p.X, Y = p.Y, u1, u2, t, t2, I, i1, i2, z_inv, den_inv, iX, iY, ench, sm1, magic, s, one;
    fe_const_sqrtm1(sm1); fe_const_invsqrt_a_minus_d(magic); fe_1(one);
    fe_add(t, p.Z, Y); fe_sub(t2, p.Z, Y); fe_mul(u1, t2, t);     // (Z+Y)(Z-Y)
    fe_mul(u2, X, Y);
    fe_sq(t, u2); fe_mul(t, u1, t);
    (void)fe_sqrt_ratio_i<F64>(I, one, t);
    fe_mul(i1, I, u1);
    fe_mul(i2, I, u2);
    fe_mul(t, i2, p.T); fe_mul(z_inv, i1, t);
    den_inv = i2;
    fe_mul(iX, X, sm1);
    fe_mul(iY, Y, sm1);
    fe_mul(ench, i1, magic);
    fe_mul(t, p.T, z_inv);
    uint32_t rotate = (uint32_t)fe_isnegative(t);
    fe_cmov(X, iY, rotate);
    fe_cmov(Y, iX, rotate);
    fe_cmov(den_inv, ench, rotate);
    fe_mul(t, X, z_inv);
    fe_cneg(Y, (uint32_t)fe_isnegative(t));          // scale <= 2
    fe_sub2(t, p.Z, Y);                              // scale 5
    fe_mul(s, t, den_inv);
    fe_cneg(s, (uint32_t)fe_isnegative(s));
    fe_tobytes_words(out, s);
}
