// base.cu -- fixed-base multiples of the Ed25519 basepoint and RFC 8032 signing on the GPU.
// These exist to synthesise benchmark / test inputs on the device (2^20..2^24 points, 2^22
// signatures would take minutes on the host): EdwardsPoint::mul_base
// (curve25519-dalek/src/edwards.rs:918-928) and the signing half of ed25519-dalek
// (src/signing.rs, src/hazmat.rs:40-99).  Variable-time table indexing: do not use with
// production secrets.
//
// Table: T[i][j] = (j+1) * 16^i * B as packed affine Niels points, i < 64, j < 8 (48 KiB), so that
// s*B = sum_i digit_i * 16^i * B needs 64 mixed additions and no doublings.
#include <algorithm>
#include <cstring>

#include "../../include/dalek_b200.h"
#include "engine.h"
#include "hash.cuh"
#include "sc.cuh"

static inline unsigned cdiv(size_t a, unsigned b) { return (unsigned)((a + b - 1) / b); }

__global__ void __launch_bounds__(64) k_build_base_table(ge_niels_packed *__restrict__ table)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 512) return;
    int i = t >> 3, j = t & 7;
    ge_p3 B, P;
    ge_p3_basepoint(B);
    ge_niels nb; ge_affine_to_niels(nb, B.X, B.Y);
    P = B;
    for (int k = 0; k < j; k++) ge_madd(P, P, nb, 0);           // (j+1) B
    if (i) ge_mul_by_pow_2(P, P, 4 * i);                        // * 16^i
    fe zi, x, y;
    fe_invert(zi, P.Z);
    fe_mul(x, P.X, zi); fe_mul(y, P.Y, zi);
    ge_niels n; ge_affine_to_niels(n, x, y);
    ge_niels_packed pk; ge_niels_pack(pk, n);
    table[t] = pk;
}

__device__ __forceinline__ void mul_base(ge_p3 &acc, const uint32_t s_in[8], const ge_niels_packed *__restrict__ table)
{
    // reduce mod l first (B has order l), which also guarantees the radix-16 recoding fits
    uint32_t s[8];
    sc_reduce256(s, s_in);
    ge_p3_identity(acc);
    int carry = 0;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        int d = (int)((s[i >> 3] >> (4 * (i & 7))) & 15) + carry;
        carry = (d + 8) >> 4;
        d -= carry << 4;
        if (d != 0) {
            int a = d < 0 ? -d : d;
            const uint4 *src = reinterpret_cast<const uint4 *>(table + (i * 8 + a - 1));
            ge_niels_packed pk;
#pragma unroll
            for (int q = 0; q < 6; q++) { uint4 v = __ldg(src + q); pk.w[4 * q] = v.x; pk.w[4 * q + 1] = v.y; pk.w[4 * q + 2] = v.z; pk.w[4 * q + 3] = v.w; }
            ge_niels n; ge_niels_unpack(n, pk);
            ge_madd(acc, acc, n, (uint32_t)(d < 0));
        }
    }
}

__global__ void __launch_bounds__(128)
k_mul_base(const uint32_t *__restrict__ scalars, size_t n, const ge_niels_packed *__restrict__ table,
           uint64_t *__restrict__ out_limbs, uint32_t *__restrict__ out_comp)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = scalars[8 * i + k];
    ge_p3 P;
    mul_base(P, s, table);
    if (out_limbs) {
        uint64_t *o = out_limbs + 20 * i;
        fe_to_limbs51(o, P.X); fe_to_limbs51(o + 5, P.Y); fe_to_limbs51(o + 10, P.Z); fe_to_limbs51(o + 15, P.T);
    }
    if (out_comp) {
        uint32_t c[8]; ge_compress(c, P);
#pragma unroll
        for (int k = 0; k < 8; k++) out_comp[8 * i + k] = c[k];
    }
}

// RFC 8032 5.1.5 / 5.1.6 (ed25519-dalek src/hazmat.rs:40-99, src/signing.rs): one thread per message
__global__ void __launch_bounds__(128)
k_sign(const uint32_t *__restrict__ seeds, const uint8_t *__restrict__ msgs, const uint64_t *__restrict__ offs, size_t n,
       const ge_niels_packed *__restrict__ table, uint32_t *__restrict__ pks, uint32_t *__restrict__ sigs)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t seed[8], hd[16], a[8], prefix[8];
#pragma unroll
    for (int k = 0; k < 8; k++) seed[k] = seeds[8 * i + k];
    sha512_state st;
    sha512_init(st); sha512_update_words(st, seed); sha512_final_words(st, hd);
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = hd[k]; prefix[k] = hd[8 + k]; }
    a[0] &= 0xfffffff8u; a[7] &= 0x3fffffffu; a[7] |= 0x40000000u;          // clamp
    ge_p3 P;
    mul_base(P, a, table);
    uint32_t A[8]; ge_compress(A, P);
    const uint8_t *m = msgs + offs[i];
    size_t len = (size_t)(offs[i + 1] - offs[i]);
    sha512_init(st); sha512_update_words(st, prefix); sha512_update(st, m, len); sha512_final_words(st, hd);
    uint32_t r[8]; sc_reduce512(r, hd);
    mul_base(P, r, table);
    uint32_t R[8]; ge_compress(R, P);
    sha512_init(st); sha512_update_words(st, R); sha512_update_words(st, A); sha512_update(st, m, len); sha512_final_words(st, hd);
    uint32_t k_[8], ka[8], ar[8], S[8];
    sc_reduce512(k_, hd);
    sc_reduce256(ar, a);
    sc_mul(ka, k_, ar);
    sc_add(S, ka, r);
#pragma unroll
    for (int k = 0; k < 8; k++) { pks[8 * i + k] = A[k]; sigs[16 * i + k] = R[k]; sigs[16 * i + 8 + k] = S[k]; }
}

int base_table_ensure(dalek_b200_ctx *ctx)
{
    if (ctx->base_table_ready) return 0;
    int rc;
    if ((rc = ws_reserve(ctx, ctx->base_table, 512 * sizeof(ge_niels_packed)))) return rc;
    k_build_base_table<<<8, 64, 0, ctx->stream>>>((ge_niels_packed *)ctx->base_table.p);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->base_table_ready = true;
    return 0;
}

extern "C" {

int dalek_b200_edwards_mul_base_batch(dalek_b200_ctx *ctx, const uint8_t *scalars, size_t n, uint64_t *out_limbs,
                                      uint8_t *out_compressed)
{
    if (!ctx || (n && (!scalars || (!out_limbs && !out_compressed)))) return DALEK_E_INVALID_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    cudaStream_t st = ctx->stream;
    if ((rc = base_table_ensure(ctx))) return rc;
    if (!n) return 0;
    if ((rc = ws_reserve(ctx, ctx->scalars, n * 32))) return rc;
    if (out_limbs && (rc = ws_reserve(ctx, ctx->points_in, n * 160))) return rc;
    if (out_compressed && (rc = ws_reserve(ctx, ctx->misc1, n * 32))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    k_mul_base<<<cdiv(n, 128), 128, 0, st>>>((const uint32_t *)ctx->scalars.p, n, (const ge_niels_packed *)ctx->base_table.p,
                                             out_limbs ? (uint64_t *)ctx->points_in.p : nullptr,
                                             out_compressed ? (uint32_t *)ctx->misc1.p : nullptr);
    ctx->launches++;
    if (out_limbs) CUDA_TRY(ctx, cudaMemcpyAsync(out_limbs, ctx->points_in.p, n * 160, cudaMemcpyDeviceToHost, st));
    if (out_compressed) CUDA_TRY(ctx, cudaMemcpyAsync(out_compressed, ctx->misc1.p, n * 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

int ed25519_b200_sign_batch_flat(dalek_b200_ctx *ctx, const uint8_t *seeds, const uint8_t *msgs_flat,
                                 const uint64_t *msg_offsets, size_t n, uint8_t *pubkeys_out, uint8_t *sigs_out)
{
    if (!ctx || (n && (!seeds || !msg_offsets || !pubkeys_out || !sigs_out))) return DALEK_E_INVALID_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    cudaStream_t st = ctx->stream;
    if ((rc = base_table_ensure(ctx))) return rc;
    if (!n) return 0;
    for (size_t i = 0; i < n; i++) if (msg_offsets[i] > msg_offsets[i + 1]) return DALEK_E_INVALID_ARG;   // offsets must not decrease
    size_t mbytes = (size_t)msg_offsets[n];
    if ((rc = ws_reserve(ctx, ctx->scalars, n * 32))) return rc;
    if ((rc = ws_reserve(ctx, ctx->misc1, mbytes + 16))) return rc;
    if ((rc = ws_reserve(ctx, ctx->misc2, (n + 1) * 8))) return rc;
    if ((rc = ws_reserve(ctx, ctx->points_in, n * 96))) return rc;
    uint32_t *d_pk = (uint32_t *)ctx->points_in.p, *d_sig = d_pk + n * 8;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, seeds, n * 32, cudaMemcpyHostToDevice, st));
    if (mbytes) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->misc1.p, msgs_flat, mbytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->misc2.p, msg_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    k_sign<<<cdiv(n, 128), 128, 0, st>>>((const uint32_t *)ctx->scalars.p, (const uint8_t *)ctx->misc1.p,
                                         (const uint64_t *)ctx->misc2.p, n, (const ge_niels_packed *)ctx->base_table.p, d_pk, d_sig);
    ctx->launches++;
    CUDA_TRY(ctx, cudaMemcpyAsync(pubkeys_out, d_pk, n * 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(sigs_out, d_sig, n * 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
