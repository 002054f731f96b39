// straus_vt.cu -- variable-time Straus for SMALL inputs: what EdwardsPoint::optional_multiscalar_mul dispatches to
// below 190 points (curve25519-dalek/src/edwards.rs:1025-1029 -> backend/serial/scalar_mul/straus.rs:159-200).
//
// The bucket pipeline (msm.cu) costs ~27 dependent kernel launches whatever n is; for a few dozen points the call is
// then launch- and latency-bound.  This path is three launches:
//
//   k_straus_prepare   one thread per (scalar, point): width-5 non-adjacent form of the scalar
//                      (Scalar::non_adjacent_form, scalar.rs:955-1007) and the table [A, 3A, ..., 15A] of projective
//                      Niels points (NafLookupTable5::from, window.rs:201-211)
//   k_straus_vartime   one warp per 8 points, one 4-lane group per point (warp4_f64.cuh): for i = 255..0
//                      Q <- 2Q; Q <- Q +/- table[|naf_i| / 2] when naf_i != 0 (straus.rs:181-197, window.rs:187-192);
//                      the reference shares ONE accumulator between the points of a call, here every group keeps its
//                      own (the doublings of different groups run in parallel lanes) and the warp adds its eight
//                      accumulators at the end -- the same group element
//   k_combine (msm.cu) sum of the per-warp results, encoding
//
// Variable time by contract (VartimeMultiscalarMul): digits select table entries by address and skip additions.
#include "../../include/dalek_b200.h"
#include "engine.h"
#include "warp4.cuh"
#include "warp4_f64.cuh"

// Scalar::non_adjacent_form(5), scalar.rs:955-1007, on four 64-bit words.  The reference requires bit 255 clear
// (debug_assert, scalar.rs:960) and produces 256 digits; this boundary takes any 256-bit value, so NAF_LEN digits are
// produced: for reference-legal scalars the digits beyond 255 are zero and the first 256 are the reference's.
#define NAF_LEN 264
__device__ __forceinline__ void naf5(int8_t *__restrict__ naf /* NAF_LEN */, const uint32_t s[8])
{
    uint64_t x[6];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = (uint64_t)s[2 * i] | ((uint64_t)s[2 * i + 1] << 32);
    x[4] = 0; x[5] = 0;
    const uint64_t width = 32, window_mask = 31;
    uint32_t pos = 0;
    uint64_t carry = 0;
    for (int i = 0; i < NAF_LEN; i++) naf[i] = 0;
    while (pos < NAF_LEN - 5) {
        const uint32_t idx = pos >> 6, bit = pos & 63;
        uint64_t bit_buf;
        if (bit < 64 - 5) bit_buf = x[idx] >> bit;
        else bit_buf = (x[idx] >> bit) | (x[idx + 1] << (64 - bit));
        const uint64_t window = carry + (bit_buf & window_mask);
        if ((window & 1) == 0) { pos += 1; continue; }              // scalar.rs:990-996
        if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)((int64_t)window - (int64_t)width); }
        pos += 5;
    }
}

__device__ __forceinline__ void store_pniels(ge_pniels_packed *dst, const ge64_p3 &p, const fe64 &d2)
{
    // EdwardsPoint::as_projective_niels (edwards.rs:528-535), canonical 32-byte coordinates
    fe64 ypx, ymx, t2d;
    fe64_add(ypx, p.Y, p.X); fe64_sub(ymx, p.Y, p.X); fe64_mul(t2d, p.T, d2);
    ge_pniels n;
    fe64_to_fe(n.YpX, ypx); fe64_to_fe(n.YmX, ymx); fe64_to_fe(n.Z, p.Z); fe64_to_fe(n.T2d, t2d);
    ge_pniels_packed pk; ge_pniels_pack(pk, n);
    uint4 *o = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = make_uint4(pk.w[4 * k], pk.w[4 * k + 1], pk.w[4 * k + 2], pk.w[4 * k + 3]);
}

template <int KIND>
__global__ void __launch_bounds__(64)
k_straus_prepare(const uint32_t *__restrict__ scalars, const void *__restrict__ points, size_t n, int8_t *__restrict__ nafs,
                 ge_pniels_packed *__restrict__ tables)
{
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t s[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = scalars[8 * j + k];
    naf5(nafs + NAF_LEN * j, s);
    fe64 d2; { fe k; fe_const_2d(k); fe64_from_fe_limbs(d2, k); }
    // A as an extended point: identity + the prepared (Niels / projective Niels) form of the input
    ge64_p3 A, A2, acc;
    ge64_identity(A);
    if (KIND == PK_NIELS) {
        ge_niels_packed pk = reinterpret_cast<const ge_niels_packed *>(points)[j];
        ge64_niels nl; ge64_niels_unpack(nl, pk);
        ge64_madd(A, A, nl, 0u);
    } else {
        ge_pniels_packed pk = reinterpret_cast<const ge_pniels_packed *>(points)[j];
        ge64_pniels pn; ge64_pniels_unpack(pn, pk);
        ge64_padd(A, A, pn, 0u);
    }
    ge64_dbl(A2, A);                                               // window.rs:206
    ge64_pniels A2n;
    fe64_add(A2n.YpX, A2.Y, A2.X); fe64_sub(A2n.YmX, A2.Y, A2.X); A2n.Z = A2.Z; fe64_mul(A2n.T2d, A2.T, d2);
    acc = A;
    ge_pniels_packed *tab = tables + 8 * j;
    store_pniels(tab, acc, d2);                                    // Ai[0] = A
#pragma unroll 1
    for (int i = 1; i < 8; i++) {                                  // Ai[i] = A2 + Ai[i-1]  (window.rs:207-209)
        ge64_padd(acc, acc, A2n, 0u);
        store_pniels(tab + i, acc, d2);
    }
}

// p <- p + q or p - q for a projective Niels table entry (curve_models.rs:411-452 + :365-372), four lanes
__device__ __forceinline__ void w4f_padd(w4f_point &p, const ge_pniels_packed &pk, uint32_t neg, uint32_t role)
{
    ge64_pniels q; ge64_pniels_unpack(q, pk);                      // coordinates in [0, 2^51): scale 2
    fe64 qp = q.YpX, qm = q.YmX;
    { fe64 t = qp; fe64_cmov(qp, qm, neg); fe64_cmov(qm, t, neg); }
    fe64 A, B, f, g, r, a, b, c, zz, D, E, H, DpC, DmC, F, G;
    fe64_sub(A, p.Y, p.X); fe64_add(B, p.Y, p.X);                  // 2, 2
    fe64_sel4(f, A, B, p.T, p.Z, role);
    fe64_sel4(g, qm, qp, q.T2d, q.Z, role);
    fe64_mul(r, f, g);                                             // <= 2 x 2
    fe64_gbcast(a, r, 0); fe64_gbcast(b, r, 1); fe64_gbcast(c, r, 2); fe64_gbcast(zz, r, 3);
    fe64_add(D, zz, zz);
    fe64_sub(E, b, a); fe64_add(H, b, a);
    fe64_add(DpC, D, c); fe64_sub(DmC, D, c);
    fe64_carry(DmC, DmC);
    F = DmC; fe64_cmov(F, DpC, neg);                               // T of the completed point
    G = DpC; fe64_cmov(G, DmC, neg);                               // Z of the completed point
    fe64_sel4(f, F, G, DmC, E, role);                              // X3 = F E, Y3 = G H, Z3 = DmC DpC, T3 = E H
    fe64_sel4(g, E, H, DpC, H, role);
    fe64_mul(r, f, g);
    fe64_gbcast(p.X, r, 0); fe64_gbcast(p.Y, r, 1); fe64_gbcast(p.Z, r, 2); fe64_gbcast(p.T, r, 3);
}

__device__ __forceinline__ void w4f_shfl_down(w4f_point &o, const w4f_point &p, int delta_lanes)
{
    const fe64 *src[4] = {&p.X, &p.Y, &p.Z, &p.T};
    fe64 *dst[4] = {&o.X, &o.Y, &o.Z, &o.T};
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const long long v = __double_as_longlong(src[c]->v[k]);
            const uint32_t lo = __shfl_down_sync(0xffffffffu, (uint32_t)v, delta_lanes), hi = __shfl_down_sync(0xffffffffu, (uint32_t)(v >> 32), delta_lanes);
            dst[c]->v[k] = __longlong_as_double((long long)(((uint64_t)hi << 32) | lo));
        }
}

__global__ void __launch_bounds__(32)
k_straus_vartime(const int8_t *__restrict__ nafs, const ge_pniels_packed *__restrict__ tables, size_t n, ge_p3_raw *__restrict__ partial)
{
    const uint32_t lane = threadIdx.x, role = lane & 3, grp = lane >> 2;
    const size_t j = (size_t)blockIdx.x * 8 + grp;
    const bool live = j < n;
    const int8_t *naf = nafs + NAF_LEN * (live ? j : 0);
    const ge_pniels_packed *tab = tables + 8 * (live ? j : 0);
    fe64 d2; { fe k; fe_const_2d(k); fe64_from_fe_limbs(d2, k); }
    w4f_point Q;
    w4f_identity(Q);
    // leading zero digits: doubling the identity changes nothing (straus.rs:181-190 starts from the identity)
    int top = -1;
    for (int i = NAF_LEN - 1; i >= 0; i--) {
        const int d = live ? naf[i] : 0;
        if (__any_sync(0xffffffffu, d != 0)) { top = i; break; }
    }
#pragma unroll 1
    for (int i = top; i >= 0; i--) {
        const int d = live ? naf[i] : 0;
        const bool any = __any_sync(0xffffffffu, d != 0);
        if (i != top) w4f_dbl(Q, role, any);
        if (any) {                                                 // uniform per warp: full-mask shuffles inside
            const uint32_t neg = d < 0, e = (uint32_t)(neg ? -d : d) >> 1;      // window.rs:187-192: entry |x| / 2
            ge_pniels_packed pk;
            const uint4 *src = reinterpret_cast<const uint4 *>(tab + (d ? e : 0));
#pragma unroll
            for (int k = 0; k < 8; k++) { uint4 v = src[k]; pk.w[4 * k] = v.x; pk.w[4 * k + 1] = v.y; pk.w[4 * k + 2] = v.z; pk.w[4 * k + 3] = v.w; }
            w4f_point R = Q;
            w4f_padd(R, pk, neg, role);
            if (d != 0) Q = R;                                     // groups with a zero digit keep Q
        }
    }
    // the warp's eight accumulators -> one point
    for (int delta = 16; delta >= 4; delta >>= 1) {
        w4f_point X; w4f_shfl_down(X, Q, delta);
        w4f_add(Q, X, d2, role);
    }
    if (grp == 0) {
        ge_p3 o; w4f_to_p3(o, Q);
        fe mine; fe_sel4(mine, o.X, o.Y, o.Z, o.T, role);
        uint32_t *dst = partial[blockIdx.x].w + 10 * role;
#pragma unroll
        for (int i = 0; i < 10; i += 2) *reinterpret_cast<uint2 *>(dst + i) = make_uint2(mine.v[i], mine.v[i + 1]);
    }
}

// sum scalars[i] * points[i] for n < 2^16 prepared points (PK_NIELS / PK_PNIELS); result like msm_full
int straus_vartime_msm(dalek_b200_ctx *ctx, const uint32_t *d_scalars, const void *d_points, int point_kind, size_t n,
                       MsmResult *d_result)
{
    int rc;
    cudaStream_t st = ctx->stream;
    const size_t n1 = n ? n : 1, nwarps = (n + 7) / 8;
    if ((rc = ws_reserve(ctx, ctx->digits, n1 * NAF_LEN))) return rc;
    if ((rc = ws_reserve(ctx, ctx->red_a, n1 * 8 * sizeof(ge_pniels_packed)))) return rc;
    if ((rc = ws_reserve(ctx, ctx->red_b, (nwarps ? nwarps : 1) * sizeof(ge_p3_raw)))) return rc;
    if (n) {
        const unsigned grid = (unsigned)((n + 63) / 64);
        if (point_kind == PK_NIELS)
            k_straus_prepare<PK_NIELS><<<grid, 64, 0, st>>>(d_scalars, d_points, n, (int8_t *)ctx->digits.p, (ge_pniels_packed *)ctx->red_a.p);
        else
            k_straus_prepare<PK_PNIELS><<<grid, 64, 0, st>>>(d_scalars, d_points, n, (int8_t *)ctx->digits.p, (ge_pniels_packed *)ctx->red_a.p);
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st));
        k_straus_vartime<<<(unsigned)nwarps, 32, 0, st>>>((const int8_t *)ctx->digits.p, (const ge_pniels_packed *)ctx->red_a.p, n, (ge_p3_raw *)ctx->red_b.p);
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st));
        ctx->last_kernel_launches = 1;
        ctx->launches += 2;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return msm_combine_windows(ctx, (const ge_p3_raw *)ctx->red_b.p, (int)nwarps, 1, 1, d_result);   // nwin = 1: plain sum of the warps' results
}
