// single.cu -- many independent Ed25519 verifications, one verdict per signature:
//   VerifyingKey::from_bytes + verify        ed25519-dalek/src/verifying.rs:167-175, :203-219
//   VerifyingKey::verify_strict              ed25519-dalek/src/verifying.rs:359-382
//   RCompute                                 ed25519-dalek/src/verifying.rs:496-557
//   EdwardsPoint::vartime_double_scalar_mul_basepoint   curve25519-dalek/src/edwards.rs:1388-1397
//       (serial backend scalar_mul/vartime_double_base.rs:23-72)
// Unlike verify_batch, `verify` recomputes R' = [s]B - [k]A and compares its ENCODING with the signature's R
// bytes, so a non-canonical R fails here although it passes the batch equation (ed25519-dalek README,
// "Validation criteria"); VALIDATIONVECTORS pins both behaviours (tests).
//
// Two paths, same results.  When the keys of a call repeat, the doublings are paid once per KEY (per-key comb tables,
// k_each_key_* and k_verify_each_comb below: 128 mixed additions per signature, no doubling).  Otherwise (plain kernel,
// k_verify_each): one thread per signature.  [s]B + [k](-A) is computed on the FP64 field with radix-16 signed digits
// (scalar.rs:1019-1051): 63 x 4 doublings, 64 mixed additions from the shared 8-entry table of B and 64 additions
// from the thread's own 8-entry table of -A (local memory).  The reference interleaves width-5 / width-8 NAFs;
// fixed radix-16 keeps the lanes of a warp on the same schedule.  Same group element, hence the same encoding.
#include <algorithm>
#include <cstring>

#include "../../include/dalek_b200.h"
#include "engine.h"
#include "ge64.cuh"
#include "hash.cuh"
#include "sc.cuh"

static inline unsigned cdiv(size_t a, unsigned b) { return (unsigned)((a + b - 1) / b); }

__device__ __forceinline__ void radix16(int8_t d[64], const uint32_t w[8])
{
    int carry = 0;
#pragma unroll 1
    for (int pos = 0; pos < 64; pos++) {
        int v = (int)((w[pos >> 3] >> (4 * (pos & 7))) & 15) + carry;
        if (pos < 63) { carry = (v + 8) >> 4; v -= carry << 4; }
        d[pos] = (int8_t)v;
    }
}

// [8]P == identity  (EdwardsPoint::is_small_order, C/edwards.rs:1405-1407)
__device__ __forceinline__ uint32_t is_small_order(const ge_p3 &p)
{
    ge_p3 q;
    ge_mul_by_pow_2(q, p, 3);
    return ge_is_identity(q);
}

#ifndef EACH_MIN_BLOCKS
#define EACH_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(128, EACH_MIN_BLOCKS)
k_verify_each(const uint8_t *__restrict__ msgs, const uint64_t *__restrict__ offs, const uint32_t *__restrict__ sigs,
              const uint32_t *__restrict__ keys, size_t n, int strict, const ge_niels_packed *__restrict__ base_row0,
              uint8_t *__restrict__ out)
{
    __shared__ double s_B[8 * 15];                               // (j+1) B as balanced FP64 affine Niels, j = 0..7
    if (threadIdx.x < 8) {
        ge64_niels e; ge64_niels_unpack(e, base_row0[threadIdx.x]);
        fe64 c;
        double *dst = s_B + 15 * threadIdx.x;
        fe64_carry(c, e.ypx);  for (int k = 0; k < 5; k++) dst[k] = c.v[k];
        fe64_carry(c, e.ymx);  for (int k = 0; k < 5; k++) dst[5 + k] = c.v[k];
        fe64_carry(c, e.xy2d); for (int k = 0; k < 5; k++) dst[10 + k] = c.v[k];
    }
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t R[8], s[8], Ak[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { R[k] = sigs[16 * i + k]; s[k] = sigs[16 * i + 8 + k]; Ak[k] = keys[8 * i + k]; }
    // k = SHA-512(R || A || M) mod l  (verifying.rs:515-523)
    uint32_t h[8];
    {
        uint32_t dig[16];
        const uint64_t lo = offs[i], hi = offs[i + 1];
        sha512_ram(dig, R, Ak, msgs + lo, (size_t)(hi - lo));
        sc_reduce512(h, dig);
    }
    ge_p3 A;
    fe_1(A.Z);
    const uint32_t okA = ge_decompress_affine<1>(A.X, A.Y, Ak);               // verifying.rs:167-175
    if (!okA) { fe_0(A.X); fe_1(A.Y); }
    fe_mul(A.T, A.X, A.Y);
    const uint32_t okS = sc_is_canonical(s);                                  // signature.rs:89-94, :149-169
    uint32_t okR = 1, small = 0;
    if (strict) {                                                             // verifying.rs:366-376
        ge_p3 Rp;
        fe_1(Rp.Z);
        okR = ge_decompress_affine<1>(Rp.X, Rp.Y, R);
        if (!okR) { fe_0(Rp.X); fe_1(Rp.Y); }
        fe_mul(Rp.T, Rp.X, Rp.Y);
        small = is_small_order(Rp) | is_small_order(A);
    }
    if (!okS) {                                                               // keep the digits in range; the verdict is fixed below
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = 0;
    }
    // table of (j+1) * (-A), j = 0..7
    ge64_pniels T[8];
    {
        ge_p3 nA = A, P;
        fe_neg(nA.X, A.X); fe_carry(nA.X, nA.X);
        fe_neg(nA.T, A.T); fe_carry(nA.T, nA.T);
        ge_pniels pn1, pn; ge_p3_to_pniels(pn1, nA);
        P = nA;
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            if (j) ge_padd(P, P, pn1, 0);
            ge_p3_to_pniels(pn, P);
            fe64_from_fe(T[j].YpX, pn.YpX); fe64_from_fe(T[j].YmX, pn.YmX); fe64_from_fe(T[j].Z, pn.Z); fe64_from_fe(T[j].T2d, pn.T2d);
        }
    }
    int8_t ds[64], dk[64];
    radix16(ds, s);
    radix16(dk, h);
    ge64_p3 acc; ge64_identity(acc);
#pragma unroll 1
    for (int pos = 63; pos >= 0; pos--) {
        if (pos != 63) { ge64_dbl(acc, acc); ge64_dbl(acc, acc); ge64_dbl(acc, acc); ge64_dbl(acc, acc); }
        const int a = ds[pos], b = dk[pos];
        if (a) {
            const int m = a < 0 ? -a : a;
            ge64_niels q;
            const double *row = s_B + 15 * (m - 1);
#pragma unroll
            for (int k = 0; k < 5; k++) { q.ypx.v[k] = row[k]; q.ymx.v[k] = row[5 + k]; q.xy2d.v[k] = row[10 + k]; }
            ge64_madd(acc, acc, q, (uint32_t)(a < 0));
        }
        if (b) {
            const int m = b < 0 ? -b : b;
            ge64_padd(acc, acc, T[m - 1], (uint32_t)(b < 0));
        }
    }
    ge_p3 Rc; ge64_to_p3(Rc, acc);
    uint32_t enc[8];
    ge_compress<1>(enc, Rc);                                                  // RCompute::finish, verifying.rs:553-556
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) diff |= enc[k] ^ R[k];
    // precedence: key decoding happens in from_bytes, before verify can be called; then the s check of
    // Signature parsing; then everything else is a Verify error
    uint8_t v = ED25519_ERR_VERIFY;
    if (!okA) v = ED25519_ERR_POINT_DECOMPRESSION;
    else if (!okS) v = ED25519_ERR_SCALAR_FORMAT;
    else if (okR && !small && diff == 0) v = DALEK_OK;
    out[i] = v;
}

static int verify_each_dev(dalek_b200_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_offs, const uint32_t *d_sigs,
                           const uint32_t *d_keys, size_t n, int strict, uint8_t *d_out, cudaStream_t st)
{
    if (!n) return 0;
    k_verify_each<<<cdiv(n, 128), 128, 0, st>>>(d_msgs, d_offs, d_sigs, d_keys, n, strict, (const ge_niels_packed *)ctx->base_table.p, d_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// ---- keys that sign many signatures: per-key comb tables ----------------------------------------------------------
// R' = [s]B - [k]A costs 252 doublings because A is only known when the call arrives.  When the batch holds few distinct
// keys (a validator set, an exchange's hot keys: bench.py's 2^22 signatures come from 1024 keys), the doublings can be
// paid once per KEY: with the 64 x 8 multiples (j+1) 16^i A of every distinct key in device memory (the fixed-base table
// of B has the same shape), [s]B - [k]A = sum_i s_i (16^i B) - k_i (16^i A) over radix-16 signed digits is 128 mixed
// additions and no doubling -- the comb of k_double_base_comb (straus.cu) with one table gathered per signature.  Same
// group element as the reference's vartime_double_scalar_mul_basepoint, hence the same encoding and verdict; keys are
// de-duplicated, decompressed and tabulated once per call (16 KiB x 4 = 64 KiB of table per key: 1024 keys stay in L2).
#define EACH_K 4                        // signatures per thread of the comb kernel: one shared inversion
#define EACH_ENT 16                     // doubles per table entry: y+x | y-x | 2dxy as balanced limbs, padded to 128 bytes
#define EACH_KEY_DOUBLES (64 * 8 * EACH_ENT)

// one thread per distinct key: decompress, 16^i A for i = 0..63 (63 x 4 doublings), small-order mark
__global__ void __launch_bounds__(64)
k_each_key_pow16(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ uniq, size_t nkeys, ge_p3_raw *__restrict__ pw,
                 uint8_t *__restrict__ kstat)
{
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nkeys) return;
    const size_t i = uniq[slot];
    uint32_t Ak[8];
#pragma unroll
    for (int k = 0; k < 8; k++) Ak[k] = keys[8 * i + k];
    ge_p3 A;
    fe_1(A.Z);
    const uint32_t ok = ge_decompress_affine<1>(A.X, A.Y, Ak);             // verifying.rs:167-175
    if (!ok) { fe_0(A.X); fe_1(A.Y); }
    fe_mul(A.T, A.X, A.Y);
    kstat[slot] = (uint8_t)((ok ? 0 : 1) | (is_small_order(A) ? 2 : 0));
    ge64_p3 P; ge64_from_p3(P, A);
#pragma unroll 1
    for (int pos = 0; pos < 64; pos++) {
        if (pos) { ge64_dbl(P, P); ge64_dbl(P, P); ge64_dbl(P, P); ge64_dbl(P, P); }
        ge_p3 q; ge64_to_p3(q, P);
        ge_p3_raw r; ge_p3_store_raw(r, q);
        uint4 *o = reinterpret_cast<uint4 *>(pw + slot * 64 + pos);
#pragma unroll
        for (int w = 0; w < 10; w++) o[w] = make_uint4(r.w[4 * w], r.w[4 * w + 1], r.w[4 * w + 2], r.w[4 * w + 3]);
    }
}

// one thread per table entry: (j+1) 16^i A as affine Niels coordinates in balanced FP64 limbs
__global__ void __launch_bounds__(128)
k_each_key_rows(const ge_p3_raw *__restrict__ pw, size_t nkeys, double *__restrict__ tab)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nkeys * 512) return;
    const int j = (int)(t & 7);
    ge_p3 P;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(pw + (t >> 3));
        ge_p3_raw r;
#pragma unroll
        for (int w = 0; w < 10; w++) { uint4 v = src[w]; r.w[4 * w] = v.x; r.w[4 * w + 1] = v.y; r.w[4 * w + 2] = v.z; r.w[4 * w + 3] = v.w; }
        ge_p3_load_raw(P, r);
    }
    ge_pniels nb; ge_p3_to_pniels(nb, P);
    ge_p3 Q = P;
#pragma unroll 1
    for (int k = 0; k < j; k++) ge_padd(Q, Q, nb, 0);                      // (j+1) * 16^i A
    fe zi, x, y;
    fe_invert_f64(zi, Q.Z);
    fe_mul(x, Q.X, zi); fe_mul(y, Q.Y, zi);
    ge_niels nl; ge_affine_to_niels(nl, x, y);
    fe64 e[3];
    fe64_from_fe(e[0], nl.ypx); fe64_from_fe(e[1], nl.ymx); fe64_from_fe(e[2], nl.xy2d);
    double *dst = tab + t * EACH_ENT;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int k = 0; k < 5; k++) dst[5 * c + k] = e[c].v[k];
    dst[15] = 0.0;
}

// The y coordinates of the eight points of small order (identity, order 2, order 4 twice, order 8 four times), canonical:
// 1, -1, 0, y8, -y8 (constants::EIGHT_TORSION, u64/constants.rs:196-340).  verify_strict rejects a signature whose R is one of
// them (verifying.rs:366-376); once the encodings of R' and R agree, R's bytes are canonical and a comparison of bytes decides.
__constant__ uint32_t c_small_y[5][8] = {
    {0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0xffffffecu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7fffffffu},
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x706a17c7u, 0x4fd84d3du, 0x760b3cbau, 0x0f67100du, 0xfa53202au, 0xc6cc392cu, 0x77fdc74eu, 0x7a03ac92u},
    {0x8f95e826u, 0xb027b2c2u, 0x89f4c345u, 0xf098eff2u, 0x05acdfd5u, 0x3933c6d3u, 0x880238b1u, 0x05fc536du}};

// one thread per signature: 64 additions from B's table (shared memory) and 64 from the key's table (L2), compress, compare
__global__ void __launch_bounds__(128, 3)
k_verify_each_comb(const uint32_t *__restrict__ sigs, const uint32_t *__restrict__ hs, const uint8_t *__restrict__ bad_s,
                   const uint32_t *__restrict__ rep, const uint32_t *__restrict__ dense, const uint8_t *__restrict__ kstat,
                   const double *__restrict__ tab, const ge_niels_packed *__restrict__ base_table, size_t i0, size_t n, int strict,
                   uint8_t *__restrict__ out)
{
    extern __shared__ double s_B[];                                        // 64 x 8 x 15: (j+1) 16^i B as balanced limbs
    for (int e = threadIdx.x; e < 512; e += blockDim.x) {
        ge64_niels q; ge64_niels_unpack(q, base_table[e]);
        fe64 c;
        double *dst = s_B + 15 * e;
        fe64_carry(c, q.ypx);  for (int k = 0; k < 5; k++) dst[k] = c.v[k];
        fe64_carry(c, q.ymx);  for (int k = 0; k < 5; k++) dst[5 + k] = c.v[k];
        fe64_carry(c, q.xy2d); for (int k = 0; k < 5; k++) dst[10 + k] = c.v[k];
    }
    __syncthreads();
    // EACH_K consecutive signatures per thread: their EACH_K inversions (the x / Z, y / Z of the encodings) become one, by
    // the simultaneous-inversion pattern of field.rs:239-273 -- 265 field operations per signature become 265 / 4 + 3
    const size_t first = i0 + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * EACH_K;
    if (first >= n) return;
    fe PX[EACH_K], PY[EACH_K], PZ[EACH_K], pre[EACH_K];
    uint8_t pending[EACH_K];
#pragma unroll 1
    for (int q = 0; q < EACH_K; q++) {
        const size_t i = first + q;
        if (i >= n) { fe_0(PX[q]); fe_1(PY[q]); fe_1(PZ[q]); pending[q] = 0xff; continue; }
        uint32_t s[8], h[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { s[k] = sigs[16 * i + 8 + k]; h[k] = hs[8 * i + k]; }
        const uint32_t okS = !bad_s[i];
        if (!okS) {                                                        // keep the digits in range; the verdict is fixed below
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = 0;
        }
        const uint32_t slot = dense[rep[i]];
        const uint32_t ks = kstat[slot];
        const double2 *TA = reinterpret_cast<const double2 *>(tab + (size_t)slot * EACH_KEY_DOUBLES);
        ge64_p3 acc; ge64_identity(acc);
        int cs = 0, ch = 0;                                                // radix-16 recoding carries (scalar.rs:1040-1046)
        uint32_t ws = 0, wh = 0;
#pragma unroll 1
        for (int pos = 0; pos < 64; pos++) {
            if ((pos & 7) == 0) { ws = s[pos >> 3]; wh = h[pos >> 3]; }
            int dsg = (int)(ws & 15) + cs; ws >>= 4;
            int dhg = (int)(wh & 15) + ch; wh >>= 4;
            if (pos < 63) { cs = (dsg + 8) >> 4; dsg -= cs << 4; ch = (dhg + 8) >> 4; dhg -= ch << 4; }
            // the key's entry first: its L2 latency runs under the addition from B's table
            double2 a2[8];
            const int mh = dhg < 0 ? -dhg : dhg;
            if (mh) {
                const double2 *src = TA + ((size_t)pos * 8 + (mh - 1)) * (EACH_ENT / 2);
#pragma unroll
                for (int k = 0; k < 8; k++) a2[k] = __ldg(src + k);
            }
            if (dsg) {
                const int m = dsg < 0 ? -dsg : dsg;
                const double *row = s_B + 15 * (8 * pos + (m - 1));
                ge64_niels qn;
#pragma unroll
                for (int k = 0; k < 5; k++) { qn.ypx.v[k] = row[k]; qn.ymx.v[k] = row[5 + k]; qn.xy2d.v[k] = row[10 + k]; }
                ge64_madd(acc, acc, qn, (uint32_t)(dsg < 0));
            }
            if (mh) {
                ge64_niels qn;
                qn.ypx.v[0] = a2[0].x; qn.ypx.v[1] = a2[0].y; qn.ypx.v[2] = a2[1].x; qn.ypx.v[3] = a2[1].y; qn.ypx.v[4] = a2[2].x;
                qn.ymx.v[0] = a2[2].y; qn.ymx.v[1] = a2[3].x; qn.ymx.v[2] = a2[3].y; qn.ymx.v[3] = a2[4].x; qn.ymx.v[4] = a2[4].y;
                qn.xy2d.v[0] = a2[5].x; qn.xy2d.v[1] = a2[5].y; qn.xy2d.v[2] = a2[6].x; qn.xy2d.v[3] = a2[6].y; qn.xy2d.v[4] = a2[7].x;
                ge64_madd(acc, acc, qn, (uint32_t)(dhg > 0));               // minus [k] A: positive digits subtract
            }
        }
        ge_p3 Rc; ge64_to_p3(Rc, acc);
        PX[q] = Rc.X; PY[q] = Rc.Y; PZ[q] = Rc.Z;
        pending[q] = (uint8_t)((ks & 1) ? ED25519_ERR_POINT_DECOMPRESSION : !okS ? ED25519_ERR_SCALAR_FORMAT : ((strict && (ks & 2)) ? ED25519_ERR_VERIFY : 0));
    }
    // 1 / Z of the EACH_K points with one inversion (Z is never zero: the formulas are complete)
    fe run; fe_copy(run, PZ[0]);
#pragma unroll 1
    for (int q = 1; q < EACH_K; q++) { fe_copy(pre[q], run); fe_mul(run, run, PZ[q]); }
    fe inv; fe_invert_f64(inv, run);
#pragma unroll 1
    for (int q = EACH_K - 1; q >= 0; q--) {
        fe zi;
        if (q) { fe_mul(zi, inv, pre[q]); fe_mul(inv, inv, PZ[q]); } else fe_copy(zi, inv);
        const size_t i = first + q;
        if (i >= n) continue;
        fe x, y;
        fe_mul(x, PX[q], zi); fe_mul(y, PY[q], zi);
        uint32_t enc[8];
        fe_tobytes_words(enc, y);                                          // RCompute::finish, verifying.rs:553-556 (EdwardsPoint::compress)
        enc[7] ^= (uint32_t)fe_isnegative(x) << 31;
        uint32_t R[8], diff = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { R[k] = sigs[16 * i + k]; diff |= enc[k] ^ R[k]; }
        uint32_t small = 0;
        if (strict) {                                                      // verifying.rs:366-376: R of small order (A: per key, above)
#pragma unroll 1
            for (int t = 0; t < 5; t++) {
                uint32_t d = (R[7] & 0x7fffffffu) ^ c_small_y[t][7];
#pragma unroll
                for (int k = 0; k < 7; k++) d |= R[k] ^ c_small_y[t][k];
                small |= (uint32_t)(d == 0);
            }
        }
        out[i] = pending[q] ? pending[q] : (uint8_t)((diff == 0 && !small) ? DALEK_OK : ED25519_ERR_VERIFY);
    }
}

// Verification (verify or verify_strict) of n signatures (device inputs) through per-key comb tables; *used = 0 if the keys do not repeat
// enough (or the tables would not fit) and the caller should run k_verify_each instead.
static int verify_each_comb(dalek_b200_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_offs, const uint32_t *d_sigs,
                            const uint32_t *d_keys, size_t n, int strict, uint8_t *d_out, bool *used)
{
    *used = false;
    if (!n || !ctx->opt_each_comb || !ctx->opt_field_f64) return 0;
    int rc;
    EachFront f;
    if ((rc = verify_each_front(ctx, d_msgs, d_offs, d_sigs, d_keys, n, &f))) return rc;
    const size_t tab_bytes = f.nkeys * EACH_KEY_DOUBLES * sizeof(double);
    if (tab_bytes > ((size_t)8 << 30)) return 0;
    if (ctx->opt_each_comb == 1 && f.nkeys * 8 > n) return 0;             // a table costs about what eight plain verifications cost
    cudaStream_t st = ctx->stream;
    if ((rc = ws_reserve(ctx, ctx->each_pow, f.nkeys * 64 * sizeof(ge_p3_raw)))) return rc;
    if ((rc = ws_reserve(ctx, ctx->each_tab, tab_bytes))) return rc;
    if ((rc = ws_reserve(ctx, ctx->each_kstat, f.nkeys))) return rc;
    k_each_key_pow16<<<cdiv(f.nkeys, 64), 64, 0, st>>>(d_keys, f.uniq, f.nkeys, (ge_p3_raw *)ctx->each_pow.p, (uint8_t *)ctx->each_kstat.p);
    k_each_key_rows<<<cdiv(f.nkeys * 512, 128), 128, 0, st>>>((const ge_p3_raw *)ctx->each_pow.p, f.nkeys, (double *)ctx->each_tab.p);
    const size_t smem = 512 * 15 * sizeof(double);
    if (!ctx->each_attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_verify_each_comb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->each_attr_set = true;
    }
    k_verify_each_comb<<<cdiv((n + EACH_K - 1) / EACH_K, 128), 128, smem, st>>>(d_sigs, f.hs, f.bad_s, f.rep, f.dense, (const uint8_t *)ctx->each_kstat.p,
                                                       (const double *)ctx->each_tab.p, (const ge_niels_packed *)ctx->base_table.p, 0, n, strict, d_out);
    ctx->launches += 3;
    CUDA_TRY(ctx, cudaGetLastError());
    *used = true;
    return 0;
}

extern "C" {

int ed25519_b200_verify_each_flat_dev(dalek_b200_ctx *ctx, const void *d_msgs_flat, const void *d_msg_offsets, const void *d_sigs,
                                      const void *d_pubkeys, size_t n, int strict, uint8_t *results)
{
    if (!ctx || (n && (!d_msg_offsets || !d_sigs || !d_pubkeys || !results))) return DALEK_E_INVALID_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    int rc;
    if ((rc = base_table_ensure(ctx))) return rc;
    if ((rc = ws_reserve(ctx, ctx->misc6, std::max<size_t>(1, n)))) return rc;
    bool comb = false;
    if ((rc = verify_each_comb(ctx, (const uint8_t *)d_msgs_flat, (const uint64_t *)d_msg_offsets, (const uint32_t *)d_sigs,
                               (const uint32_t *)d_pubkeys, n, strict, (uint8_t *)ctx->misc6.p, &comb))) return rc;
    if (!comb && (rc = verify_each_dev(ctx, (const uint8_t *)d_msgs_flat, (const uint64_t *)d_msg_offsets, (const uint32_t *)d_sigs,
                                       (const uint32_t *)d_pubkeys, n, strict, (uint8_t *)ctx->misc6.p, ctx->stream))) return rc;
    if (n) CUDA_TRY(ctx, cudaMemcpyAsync(results, ctx->misc6.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint8_t any = 0;
    for (size_t i = 0; i < n; i++) any |= results[i];
    return any ? ED25519_ERR_VERIFY : DALEK_OK;
}

int ed25519_b200_verify_each_flat(dalek_b200_ctx *ctx, const uint8_t *msgs_flat, const uint64_t *msg_offsets, const uint8_t *sigs,
                                  const uint8_t *pubkeys, size_t n, int strict, uint8_t *results)
{
    if (!ctx || (n && (!msg_offsets || !sigs || !pubkeys || !results))) return DALEK_E_INVALID_ARG;
    if (n && msg_offsets[0] != 0) return DALEK_E_INVALID_ARG;
    for (size_t i = 0; i < n; i++) if (msg_offsets[i] > msg_offsets[i + 1]) return DALEK_E_INVALID_ARG;   // offsets must not decrease
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    int rc;
    const size_t mbytes = n ? (size_t)msg_offsets[n] : 0;
    if ((rc = base_table_ensure(ctx))) return rc;
    if ((rc = ws_reserve(ctx, ctx->misc1, mbytes + 16))) return rc;
    if ((rc = ws_reserve(ctx, ctx->msg_offs, (n + 1) * 8))) return rc;
    if ((rc = ws_reserve(ctx, ctx->points_in, std::max<size_t>(1, n) * 96))) return rc;
    if ((rc = ws_reserve(ctx, ctx->misc6, std::max<size_t>(1, n)))) return rc;
    uint8_t *d_msgs = (uint8_t *)ctx->misc1.p, *d_sigs = (uint8_t *)ctx->points_in.p, *d_keys = d_sigs + n * 64, *d_out = (uint8_t *)ctx->misc6.p;
    uint64_t *d_offs = (uint64_t *)ctx->msg_offs.p;
    if (ctx->opt_each_comb && ctx->opt_field_f64 && n) {
        // keys may repeat: everything crosses PCIe first (the key tables need every key), then either the comb path or,
        // when the keys turn out not to repeat, the plain kernel on the resident copies
        cudaStream_t st = ctx->stream;
        if (mbytes) CUDA_TRY(ctx, cudaMemcpyAsync(d_msgs, msgs_flat, mbytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_offs, msg_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_sigs, sigs, n * 64, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_keys, pubkeys, n * 32, cudaMemcpyHostToDevice, st));
        bool comb = false;
        if ((rc = verify_each_comb(ctx, d_msgs, d_offs, (const uint32_t *)d_sigs, (const uint32_t *)d_keys, n, strict, d_out, &comb))) return rc;
        if (!comb && (rc = verify_each_dev(ctx, d_msgs, d_offs, (const uint32_t *)d_sigs, (const uint32_t *)d_keys, n, strict, d_out, st))) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(results, d_out, n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        uint8_t any = 0;
        for (size_t i = 0; i < n; i++) any |= results[i];
        return any ? ED25519_ERR_VERIFY : DALEK_OK;
    }
    // independent per signature: pieces alternate between two streams (copy-in -> kernel -> copy-out)
    cudaStream_t ss[2] = {ctx->stream, ctx->stream2};
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    const size_t piece = n >= (1u << 17) ? (size_t)1 << 16 : std::max<size_t>(1, n);
    size_t k = 0;
    for (size_t lo = 0; lo < n; lo += piece, k++) {
        const size_t m = std::min(piece, n - lo);
        cudaStream_t st = ss[k & 1];
        const size_t m0 = (size_t)msg_offsets[lo], m1 = (size_t)msg_offsets[lo + m];
        if (m1 > m0) CUDA_TRY(ctx, cudaMemcpyAsync(d_msgs + m0, msgs_flat + m0, m1 - m0, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_offs + lo, msg_offsets + lo, (m + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_sigs + lo * 64, sigs + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_keys + lo * 32, pubkeys + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        // offsets are absolute: the kernel indexes msgs by them, so pass the bases shifted by the piece start
        if ((rc = verify_each_dev(ctx, d_msgs, d_offs + lo, (const uint32_t *)(d_sigs + lo * 64), (const uint32_t *)(d_keys + lo * 32), m, strict,
                                  d_out + lo, st))) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(results + lo, d_out + lo, m, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint8_t any = 0;
    for (size_t i = 0; i < n; i++) any |= results[i];
    return any ? ED25519_ERR_VERIFY : DALEK_OK;
}

}  // extern "C"
