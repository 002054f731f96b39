/*
 * batch.c -- ed25519 batch / single verification (and RFC 8032 signing, used only to
 * synthesise inputs).  TEST INFRASTRUCTURE (oracle).
 * Restates E/batch.rs:146-251, E/verifying.rs:167-175, :203-219, :359-382, :496-557,
 * E/signature.rs:89-94, :149-160.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static void hram(uint8_t out[64], const uint8_t R[32], const uint8_t A[32],
                 const uint8_t *msg, size_t len)
{
    sha512_ctx c; sha512_init(&c);                                /* E/batch.rs:185-189 */
    sha512_update(&c, R, 32); sha512_update(&c, A, 32); sha512_update(&c, msg, len);
    sha512_final(&c, out);
}

/* one faithful transcript over signatures [lo, hi): E/batch.rs:168-205, :219-222 */
static void draw_zs(uint8_t *zs /*(hi-lo)*16*/, const uint8_t *hrams, const uint8_t *sigs,
                    size_t lo, size_t hi)
{
    merlin_transcript t, rng;
    merlin_new(&t, (const uint8_t *)"ed25519 batch verification", 26);      /* :168 */
    for (size_t i = lo; i < hi; i++)                                         /* :195-197 */
        merlin_append_message(&t, (const uint8_t *)"hram", 4, hrams + 64 * i, 64);
    for (size_t i = lo; i < hi; i++)                                         /* :199-201 */
        merlin_append_message(&t, (const uint8_t *)"sig.s", 5, sigs + 64 * i + 32, 32);
    merlin_rng_finalize_zero(&rng, &t);                                      /* :205 */
    for (size_t i = lo; i < hi; i++) merlin_rng_fill(&rng, zs + 16 * (i - lo), 16); /* :79-83, :219-222 */
}

static int verify_batch_impl(const uint8_t *const *msgs, const size_t *msg_lens,
                             const uint8_t *sigs, const uint8_t *pubkeys,
                             size_t n, size_t chunk, uint8_t *zs_out)
{
    /* VerifyingKey::from_bytes for every key (E/verifying.rs:167-175); in Rust this happens
     * before verify_batch is callable, hence it is the first error that can surface. */
    ge_p3 *As = (ge_p3 *)malloc(sizeof(ge_p3) * (n ? n : 1));
    for (size_t i = 0; i < n; i++)
        if (!ge_decompress(&As[i], pubkeys + 32 * i)) { free(As); return ED_ERR_POINT_DECOMPRESSION; }

    size_t m = 2 * n + 1;
    uint8_t *hrams = (uint8_t *)malloc(64 * (n ? n : 1));
    uint8_t *zs = (uint8_t *)malloc(16 * (n ? n : 1));
    uint8_t *scalars = (uint8_t *)calloc(m, 32);
    ge_p3 *points = (ge_p3 *)malloc(sizeof(ge_p3) * m);
    uint8_t *present = (uint8_t *)malloc(m);
    int rc = ED_OK;

    for (size_t i = 0; i < n; i++)                                           /* :179-191 */
        hram(hrams + 64 * i, sigs + 64 * i, pubkeys + 32 * i, msgs[i], msg_lens[i]);

    if (chunk == 0 || chunk > n) chunk = n ? n : 1;
    for (size_t lo = 0; lo < n; lo += chunk) {                               /* one transcript per chunk */
        size_t hi = lo + chunk < n ? lo + chunk : n;
        draw_zs(zs + 16 * lo, hrams, sigs, lo, hi);
    }
    if (zs_out) memcpy(zs_out, zs, 16 * n);

    /* InternalSignature::try_from: s canonical (E/batch.rs:208-211, E/signature.rs:89-94) */
    for (size_t i = 0; i < n; i++)
        if (!scalar_is_canonical(sigs + 64 * i + 32)) { rc = ED_ERR_SCALAR_FORMAT; goto done; }

    {
        uint8_t bcoef[32] = {0};
        for (size_t i = 0; i < n; i++) {
            uint8_t h[32], z[32] = {0}, zs_i[32];
            scalar_from_bytes_mod_order_wide(h, hrams + 64 * i);             /* :213-216 */
            memcpy(z, zs + 16 * i, 16);                                      /* :219-222 Scalar::from(u128) */
            scalar_mul(zs_i, z, sigs + 64 * i + 32);                         /* :225-230 */
            scalar_add(bcoef, bcoef, zs_i);
            memcpy(scalars + 32 * (1 + i), z, 32);
            scalar_mul(scalars + 32 * (1 + n + i), h, z);                    /* :233 */
        }
        scalar_neg(scalars, bcoef);                                          /* :241 -B_coefficient */
    }
    ge_basepoint(&points[0]); present[0] = 1;                                /* :237 */
    for (size_t i = 0; i < n; i++) {
        present[1 + i] = (uint8_t)ge_decompress(&points[1 + i], sigs + 64 * i);   /* :235 */
        if (!present[1 + i]) ge_identity(&points[1 + i]);
        points[1 + n + i] = As[i]; present[1 + n + i] = 1;                   /* :236 */
    }
    {
        ge_p3 id;
        if (!edwards_optional_multiscalar_mul(&id, scalars, points, present, m)) rc = ED_ERR_VERIFY; /* :240-244 */
        else rc = ge_is_identity(&id) ? ED_OK : ED_ERR_VERIFY;               /* :246-250 */
    }
done:
    free(As); free(hrams); free(zs); free(scalars); free(points); free(present);
    return rc;
}

int ed25519_verify_batch(const uint8_t *const *msgs, const size_t *msg_lens,
                         const uint8_t *sigs, const uint8_t *pubkeys, size_t n, uint8_t *zs_out)
{
    return verify_batch_impl(msgs, msg_lens, sigs, pubkeys, n, 0, zs_out);
}

int ed25519_verify_batch_chunked(const uint8_t *const *msgs, const size_t *msg_lens,
                                 const uint8_t *sigs, const uint8_t *pubkeys,
                                 size_t n, size_t chunk, uint8_t *zs_out)
{
    return verify_batch_impl(msgs, msg_lens, sigs, pubkeys, n, chunk, zs_out);
}

/* RCompute::compute/finish, E/verifying.rs:496-557 */
static void rcompute(uint8_t out[32], const ge_p3 *A, const uint8_t pk[32], const uint8_t sig[64],
                     const uint8_t *msg, size_t len)
{
    uint8_t h[64], k[32];
    hram(h, sig, pk, msg, len);
    scalar_from_bytes_mod_order_wide(k, h);                                  /* Scalar::from_hash */
    ge_p3 minus_A, r;
    ge_p3_neg(&minus_A, A);
    edwards_vartime_double_scalar_mul_basepoint(&r, k, &minus_A, sig + 32);  /* :553 */
    ge_compress(out, &r);
}

/* E/verifying.rs:203-219 */
int ed25519_verify(const uint8_t *msg, size_t len, const uint8_t sig[64], const uint8_t pk[32])
{
    ge_p3 A; uint8_t expected[32];
    if (!ge_decompress(&A, pk)) return ED_ERR_POINT_DECOMPRESSION;
    if (!scalar_is_canonical(sig + 32)) return ED_ERR_SCALAR_FORMAT;
    rcompute(expected, &A, pk, sig, msg, len);
    return memcmp(expected, sig, 32) == 0 ? ED_OK : ED_ERR_VERIFY;
}

/* E/verifying.rs:359-382 */
int ed25519_verify_strict(const uint8_t *msg, size_t len, const uint8_t sig[64], const uint8_t pk[32])
{
    ge_p3 A, R; uint8_t expected[32];
    if (!ge_decompress(&A, pk)) return ED_ERR_POINT_DECOMPRESSION;
    if (!scalar_is_canonical(sig + 32)) return ED_ERR_SCALAR_FORMAT;
    if (!ge_decompress(&R, sig)) return ED_ERR_VERIFY;
    if (ge_is_small_order(&R) || ge_is_small_order(&A)) return ED_ERR_VERIFY;
    rcompute(expected, &A, pk, sig, msg, len);
    return memcmp(expected, sig, 32) == 0 ? ED_OK : ED_ERR_VERIFY;
}

/* RFC 8032 5.1.5 / E/hazmat.rs:40-99 ExpandedSecretKey, E/signing.rs (input synthesis only) */
static void expand(uint8_t a[32], uint8_t prefix[32], const uint8_t seed[32])
{
    uint8_t h[64];
    sha512(h, seed, 32);
    memcpy(a, h, 32); memcpy(prefix, h + 32, 32);
    a[0] &= 248; a[31] &= 63; a[31] |= 64;
}

void ed25519_public_key(uint8_t pk[32], const uint8_t seed[32])
{
    uint8_t a[32], prefix[32]; ge_p3 B, A;
    expand(a, prefix, seed);
    ge_basepoint(&B); ge_scalarmul(&A, a, &B);
    ge_compress(pk, &A);
}

void ed25519_sign(uint8_t sig[64], const uint8_t *msg, size_t len, const uint8_t seed[32])
{
    uint8_t a[32], prefix[32], pk[32], h[64], r[32], k[32], ka[32], a_red[32];
    ge_p3 B, P;
    expand(a, prefix, seed);
    ge_basepoint(&B); ge_scalarmul(&P, a, &B); ge_compress(pk, &P);
    sha512_ctx c; sha512_init(&c);
    sha512_update(&c, prefix, 32); sha512_update(&c, msg, len); sha512_final(&c, h);
    scalar_from_bytes_mod_order_wide(r, h);
    ge_scalarmul(&P, r, &B); ge_compress(sig, &P);
    hram(h, sig, pk, msg, len);
    scalar_from_bytes_mod_order_wide(k, h);
    scalar_reduce(a_red, a);
    scalar_mul(ka, k, a_red);
    scalar_add(sig + 32, ka, r);
}
