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
// One thread per signature.  [s]B + [k](-A) is computed on the FP64 field with radix-16 signed digits
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
    if ((rc = verify_each_dev(ctx, (const uint8_t *)d_msgs_flat, (const uint64_t *)d_msg_offsets, (const uint32_t *)d_sigs,
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
