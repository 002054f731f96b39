// warp4_f64.cuh -- the 4-lane point operations of warp4.cuh over the FP64-pipe field (fe64.cuh).
//
// The final Horner pass over the bucket windows (pippenger.rs:159) is ~250 DEPENDENT doublings: its
// latency is (instructions per doubling) x (issue interval of one lone warp).  The four independent
// field multiplications of each half of a point operation run in the four lanes of a group, as in
// warp4.cuh (the SIMT form of docs/parallel-formulas.md:51-76), but on five balanced FP64 limbs: a
// squaring is 15 split products instead of 55 IMAD.WIDE plus pre-multiplications, an element is ten
// 32-bit registers to select and shuffle instead of ten limbs with masks.  All 32 lanes of a warp must
// execute these calls together (full-mask shuffles).
#pragma once
#include "ge64.cuh"

struct w4f_point { fe64 X, Y, Z, T; };       // extended point, coordinates of scale 1, replicated in the 4 lanes

__device__ __forceinline__ void fe64_sel4(fe64 &o, const fe64 &a0, const fe64 &a1, const fe64 &a2, const fe64 &a3, uint32_t role)
{
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const double lo = role & 1 ? a1.v[i] : a0.v[i], hi = role & 1 ? a3.v[i] : a2.v[i];
        o.v[i] = role & 2 ? hi : lo;
    }
}

// value held by lane `i` of this lane's group
__device__ __forceinline__ void fe64_gbcast(fe64 &o, const fe64 &mine, int i)
{
    const int src = (int)((threadIdx.x & 28u) | (uint32_t)i);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const long long v = __double_as_longlong(mine.v[k]);
        const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
        o.v[k] = __longlong_as_double((long long)(((uint64_t)hi << 32) | lo));
    }
}

// scale-1 integer limbs (fe.cuh) -> balanced doubles of scale 1
__device__ __forceinline__ void fe64_from_fe_limbs(fe64 &h, const fe &f)
{
    fe64 t;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint64_t l = (uint64_t)f.v[2 * k] + ((uint64_t)f.v[2 * k + 1] << 26);       // < 2^51 + 2^26: exact in a double
        t.v[k] = __longlong_as_double((long long)l | FE64_E52) - FE64_TWO52;
    }
    fe64_carry(h, t);
}

__device__ __forceinline__ void w4f_identity(w4f_point &p) { fe64_0(p.X); fe64_1(p.Y); fe64_1(p.Z); fe64_0(p.T); }

__device__ __forceinline__ void w4f_load(w4f_point &p, const ge_p3_raw *src)
{
    ge_p3 q;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    ge_p3_raw r;
#pragma unroll
    for (int k = 0; k < 10; k++) { uint4 v = s[k]; r.w[4 * k] = v.x; r.w[4 * k + 1] = v.y; r.w[4 * k + 2] = v.z; r.w[4 * k + 3] = v.w; }
    ge_p3_load_raw(q, r);
    fe64_from_fe_limbs(p.X, q.X); fe64_from_fe_limbs(p.Y, q.Y); fe64_from_fe_limbs(p.Z, q.Z); fe64_from_fe_limbs(p.T, q.T);
}

// p <- 2p (curve_models.rs:381-397 + :365-372).  T is refreshed only when want_t (the doubling before an addition).
__device__ __forceinline__ void w4f_dbl(w4f_point &p, uint32_t role, bool want_t)
{
    fe64 S, in, r, XX, YY, ZZ, S2, Yp, Ym, E, F, f, g;
    fe64_add(S, p.X, p.Y); fe64_carry(S, S);                  // squaring needs scale < 2
    fe64_sel4(in, p.X, p.Y, p.Z, S, role);
    fe64_sq(r, in);
    fe64_gbcast(XX, r, 0); fe64_gbcast(YY, r, 1); fe64_gbcast(ZZ, r, 2); fe64_gbcast(S2, r, 3);
    fe64_add(Yp, YY, XX);                                     // 2
    fe64_sub(Ym, YY, XX);                                     // 2
    fe64_sub(E, S2, Yp);                                      // 3   (X+Y)^2 - Y^2 - X^2
    fe64_add(F, ZZ, ZZ); fe64_sub(F, F, Ym);                  // 4   2Z^2 - (Y^2 - X^2)
    fe64_carry(F, F);                                         // 1
    fe64_sel4(f, E, Yp, Ym, E, role);                         // X3 = E F, Y3 = Yp Ym, Z3 = Ym F, T3 = E Yp
    fe64_sel4(g, F, Ym, F, Yp, role);
    fe64_mul(r, f, g);                                        // <= 3 x 2
    fe64_gbcast(p.X, r, 0); fe64_gbcast(p.Y, r, 1); fe64_gbcast(p.Z, r, 2);
    if (want_t) fe64_gbcast(p.T, r, 3);
}

// p <- p + q, both extended (edwards.rs:795-800 = :528-535 + curve_models.rs:411-430, :365-372); d2 = 2d as fe64
__device__ __forceinline__ void w4f_add(w4f_point &p, const w4f_point &q, const fe64 &d2, uint32_t role)
{
    fe64 qYpX, qYmX, qT2d, A, B, f, g, r, a, b, c, zz, D, E, H, DpC, DmC;
    fe64_add(qYpX, q.Y, q.X);                                 // 2
    fe64_sub(qYmX, q.Y, q.X);                                 // 2
    fe64_sub(A, p.Y, p.X); fe64_add(B, p.Y, p.X);             // 2, 2
    fe64_mul(qT2d, q.T, d2);                                  // as_projective_niels (replicated in the four lanes)
    fe64_sel4(f, A, B, p.T, p.Z, role);                       // a = A qYmX, b = B qYpX, c = T qT2d, zz = Z qZ
    fe64_sel4(g, qYmX, qYpX, qT2d, q.Z, role);
    fe64_mul(r, f, g);                                        // <= 2 x 2
    fe64_gbcast(a, r, 0); fe64_gbcast(b, r, 1); fe64_gbcast(c, r, 2); fe64_gbcast(zz, r, 3);
    fe64_add(D, zz, zz);                                      // 2
    fe64_sub(E, b, a); fe64_add(H, b, a);                     // 2, 2
    fe64_add(DpC, D, c); fe64_sub(DmC, D, c);                 // 3, 3
    fe64_carry(DmC, DmC);                                     // 1   (3 x 3 would break the operand rule of DmC * DpC)
    fe64_sel4(f, DmC, DpC, DmC, E, role);                     // X3 = DmC E, Y3 = DpC H, Z3 = DmC DpC, T3 = E H
    fe64_sel4(g, E, H, DpC, H, role);
    fe64_mul(r, f, g);                                        // <= 3 x 2
    fe64_gbcast(p.X, r, 0); fe64_gbcast(p.Y, r, 1); fe64_gbcast(p.Z, r, 2); fe64_gbcast(p.T, r, 3);
}

__device__ __forceinline__ void w4f_to_p3(ge_p3 &o, const w4f_point &p)
{
    fe64_to_fe(o.X, p.X); fe64_to_fe(o.Y, p.Y); fe64_to_fe(o.Z, p.Z); fe64_to_fe(o.T, p.T);
}
